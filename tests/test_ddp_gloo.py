"""CPU, world_size 2 (gloo): the data-parallel protocol of SURVEY.md 8e / DESIGN.md section 6 -- masked counts all-reduced
BEFORE the per-rank loss, loss terms normalised by GLOBAL counts, gradients all-reduced with SUM -- reproduces the
single-process gradient of the concatenated batch, and the naive "mean of per-rank means" does not.  The arithmetic
comes from the oracle; the protocol is the one deepsvg_b200.SVGLoss / SVGTransformer implement on NCCL."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import svg_oracle as O


def _rank_loss(params, cfg, cmd, arg, counts, n_total_paths):
    """One rank's share of the global loss: sums over local elements divided by global counts."""
    out = O.forward(params, cfg, cmd, arg)
    tc, ta = cmd.long(), arg
    vis = O.visibility(tc)
    wc = (O.extended_padding(tc) * vis.unsqueeze(-1).float())[..., 1:]
    tc1, ta1 = tc[..., 1:], ta[..., 1:, :]
    wa = O.CMD_ARGS_MASK[tc1].float()
    F = torch.nn.functional
    cl, al = out["command_logits"], out["args_logits"]
    ce_c = F.cross_entropy(cl.reshape(-1, 7), tc1.reshape(-1), reduction="none").reshape(tc1.shape)
    ce_a = F.cross_entropy(al.reshape(-1, al.shape[-1]), (ta1.long() + 1).reshape(-1), reduction="none").reshape(ta1.shape)
    ce_v = F.cross_entropy(out["visibility_logits"].reshape(-1, 2), vis.reshape(-1).long(), reduction="sum")
    return 1.0 * (ce_c * wc).sum() / counts[0] + 2.0 * (ce_a * wa).sum() / counts[1] + 1.0 * ce_v / n_total_paths


def _local_counts(cmd):
    tc = cmd.long()
    wc = (O.extended_padding(tc) * O.visibility(tc).unsqueeze(-1).float())[..., 1:]
    return torch.stack([wc.sum(), O.CMD_ARGS_MASK[tc[..., 1:]].float().sum()])


def _worker(rank, world, init_file, out_file):
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = _cfg()
    params = {k: v.clone().requires_grad_(True) for k, v in O.make_params(cfg, seed=1).items()}
    cmd, arg = O.synth_batch(cfg, 4, seed=77)
    half = slice(rank * 2, rank * 2 + 2)
    c, a = cmd[half], arg[half]
    counts = _local_counts(c)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)                      # SVGLoss: global masked counts first
    loss = _rank_loss(params, cfg, c, a, counts, n_total_paths=4 * cfg.max_num_groups)
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in params.values()])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                        # SVGTransformer.backward: the one collective
    lsum = loss.detach().clone()
    dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save({"flat": flat, "loss": lsum}, out_file)
    dist.destroy_process_group()


def _cfg():
    return O.make_cfg("hierarchical", d_model=32, n_heads=4, dim_feedforward=64, dim_z=24, n_layers=1, n_layers_decode=1,
                      max_num_groups=3, max_seq_len=8, args_dim=15, use_vae=False)


def test_two_rank_gradients_equal_single_process():
    tmp = tempfile.mkdtemp()
    init_file, out_file = os.path.join(tmp, "init"), os.path.join(tmp, "out.pt")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    mp.spawn(_worker, args=(2, init_file, out_file), nprocs=2, join=True)
    got = torch.load(out_file)
    cfg = _cfg()
    params = O.make_params(cfg, seed=1)
    cmd, arg = O.synth_batch(cfg, 4, seed=77)
    _, ls, grads = O.train_step(params, cfg, cmd, arg)
    ref = torch.cat([grads[k].reshape(-1) for k in params])
    assert abs(got["loss"].item() - ls["loss"].item()) < 1e-5 * ls["loss"].item()
    assert (got["flat"] - ref).norm().item() < 1e-5 * ref.norm().item()
    # the naive protocol (per-rank means, averaged) is measurably different: the counts differ between ranks
    naive = []
    for r in range(2):
        _, _, g = O.train_step(params, cfg, cmd[2 * r:2 * r + 2], arg[2 * r:2 * r + 2])
        naive.append(torch.cat([g[k].reshape(-1) for k in params]))
    naive = 0.5 * (naive[0] + naive[1])
    assert (naive - ref).norm().item() > 1e-4 * ref.norm().item()
