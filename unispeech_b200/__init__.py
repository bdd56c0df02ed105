"""unispeech_b200 -- B200-native WavLM / UniSpeech-SAT encoder hot path (hand-written sm_100a kernels behind a C ABI)."""

__version__ = "0.1.0"
