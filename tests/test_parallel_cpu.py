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
    from unispeech_b200.engine import FlatGrads
    from unispeech_b200.parallel import all_reduce_grads, shard_batch
    from unispeech_b200.wavlm import WavLM, WavLMConfig

    cfg = O.tiny_config()
    m = WavLM(WavLMConfig(vars(cfg)))
    m.load_state_dict(O.deterministic_state_dict(cfg))
    groups = []
    for lyr in m.encoder.layers:
        a = lyr.self_attn
        groups.append([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight])
        groups.append([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias])
    seen = {id(p) for g in groups for p in g}
    groups.append([p for p in m.parameters() if id(p) not in seen])
    flat = FlatGrads(groups, torch.device("cpu"))
    flat.attach()
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
