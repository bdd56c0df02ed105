"""world_size-2 gloo test (CPU) of the data-parallel path: flat gradient buffer aliasing + the single allreduce per step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wavlm_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace

    from unispeech_b200.engine import build_flat_grads
    from unispeech_b200.parallel import OverlappedGradSync, all_reduce_grads, shard_batch
    from unispeech_b200.wavlm import WavLM, WavLMConfig

    cfg = O.tiny_config()
    m = WavLM(WavLMConfig(vars(cfg)))
    m.load_state_dict(O.deterministic_state_dict(cfg))
    flat, ranges, order = build_flat_grads(m, torch.device("cpu"))   # the layout the engine builds on the GPU
    flat.attach()
    # backward-completion order: head, layers last -> first, stem, conv stack; the ranges tile the buffer without gaps
    assert order[0] == "head" and order[1] == ("layer", cfg.encoder_layers - 1) and order[-2:] == ["stem", "conv"]
    pos = 0
    for st in order:
        assert ranges[st][0] == pos
        pos = ranges[st][1]
    assert pos == flat.flat.numel()
    # q/k/v gradients are adjacent so that the fused [3D, D] weight gradient is one GEMM output
    a0 = m.encoder.layers[0].self_attn
    D = cfg.encoder_embed_dim
    assert a0.k_proj.weight.grad.data_ptr() == a0.q_proj.weight.grad.data_ptr() + 4 * D * D
    assert a0.v_proj.bias.grad.data_ptr() == a0.q_proj.bias.grad.data_ptr() + 8 * D
    for i, p in enumerate(m.parameters()):
        p.grad.fill_(float(rank + 1) * (1 + (i % 3)))  # writes through the views into the flat buffer
    all_reduce_grads(flat.flat)
    ok = True
    for i, p in enumerate(m.parameters()):
        want = 1.5 * (1 + (i % 3))  # mean of ranks 1 and 2
        ok = ok and bool(torch.allclose(p.grad, torch.full_like(p.grad, want)))
    # ---- the bucketed exchange overlapped with the backward pass: same result, buckets launched as their stages complete
    for i, p in enumerate(m.parameters()):
        p.grad.fill_(float(rank + 1) * (2 + (i % 5)))
    eng = SimpleNamespace(flat=flat, stage_ranges=ranges, stage_order=order, grad_sync=None)
    sync = OverlappedGradSync(SimpleNamespace(_engine=eng), layers_per_bucket=1)
    ok = ok and eng.grad_sync is sync and len(sync.buckets) == cfg.encoder_layers + 1
    sync.begin()
    launched = []
    for st in order[1:]:
        if st == ("layer", 0):
            continue  # a layer the backward pass never reached (layerdrop): swept up by the next stage / finish()
        sync.stage_done(st)
        launched.append(sync._next)
    ok = ok and launched == [1, 2, 3][:cfg.encoder_layers - 1] + [cfg.encoder_layers, cfg.encoder_layers + 1][-2:]
    sync.finish()
    for i, p in enumerate(m.parameters()):
        ok = ok and bool(torch.allclose(p.grad, torch.full_like(p.grad, 1.5 * (2 + (i % 5)))))
    # zero_grad(set_to_none=True) followed by attach() gives zeroed views again
    m.zero_grad(set_to_none=True)
    flat.attach()
    ok = ok and all(float(p.grad.abs().max()) == 0.0 for p in m.parameters())
    lo, hi = shard_batch(5, rank, world)
    ok = ok and (lo, hi) == ((0, 3) if rank == 0 else (3, 5))
    out[rank] = ok
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
