"""cProfile of the host side of a few training steps (where does the enqueue time go?). Run on the GPU box."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import wavlm_oracle as O
from unispeech_b200.wavlm import WavLM, WavLMConfig

dev = torch.device("cuda:0")
cfg = O.base_config()
model = WavLM(WavLMConfig(vars(cfg)))
model.load_state_dict(O.deterministic_state_dict(cfg))
model = model.to(dev).train()
B, L_ = 16, 15 * 16000
wav = torch.randn(B, L_, device=dev)
pad = torch.zeros(B, L_, dtype=torch.bool)
R = torch.randn(B, O.num_frames(L_, cfg), cfg.encoder_embed_dim, device=dev)

def step():
    if model._engine is not None and model._engine.flat is not None:
        model.grad_buffer().zero_()
        model._engine.prepared_version = None
    x, _ = model.extract_features(wav, padding_mask=pad, mask=True)
    (x.float() * R).sum().backward()

for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
