"""Training step (forward + backward [+ gradient all-reduce]) of the hot path at the BASELINE.json configs[3]
shapes: B=8 scenes per GPU, N=50 000 points, L=80 tokens, D=132 boxes, K=256 queries; train-mode BatchNorm
(batch statistics), dropout 0 (the parity configuration).  Model = Pointnet2Backbone ->
3 x BiEncoderLayer -> 6 x BiDecoderLayer with a synthetic quadratic loss; everything else of BeaUTyDETR is out of
scope.

  python benchmarks/micro_train.py                                   # 1 GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/micro_train.py   # N GPUs

  ours        eda_b200 modules; backward = recompute with differentiable torch ops + the CUDA scatter kernels;
              gradients of all ranks meet in ONE flat all-reduce (eda_b200/ddp.py)
  reference   (rank 0 of a 1-GPU run only) the reference's own compiled `_ext` behind the same op sequence as
              pointnet2_modules.py + torch layers, attention as plain torch ops under autograd
JSON on stdout (rank 0).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import attn_cases as ac  # noqa: E402

from eda_b200 import ddp, encoder_decoder_layers as edl, synthetic  # noqa: E402
from eda_b200.backbone_module import Pointnet2Backbone  # noqa: E402
from eda_b200.pointnet2 import pointnet2_utils  # noqa: E402
from oracle import attention_oracle as ao  # noqa: E402  (baseline leg only)


from eda_b200.hotpath import HotPath  # noqa: E402


def reference_forward(model, ref_ext, pc, pos, text, text_mask, det, det_mask, query, qpos):
    """Same parameters, reference kernels: compiled reference `_ext` behind the unfused op sequence, torch layers."""
    saved = pointnet2_utils._ext
    pointnet2_utils._ext = ref_ext
    try:
        ep = model.backbone(pc)
    finally:
        pointnet2_utils._ext = saved
    vis = ep["fp2_features"].transpose(1, 2).contiguous()
    esd = {k: v for k, v in model.encoder.named_parameters()}
    esd.update({k: v for k, v in model.encoder.named_buffers()})
    v, t = ao.bi_encoder(esd, "", 3, vis, pos, None, text, text_mask, det, det_mask)
    q = query
    for d in model.decoder:
        # train-mode BatchNorm1d of the position embedding: use the layer's own torch modules for that part
        sd = {k: p for k, p in d.named_parameters()}
        sd.update({k: b for k, b in d.named_buffers()})
        pos_q = d.self_posembed.position_embedding_head(qpos.transpose(1, 2).contiguous()).transpose(1, 2)
        q2 = ao.mha(sd, "self_attn.", q + pos_q, q + pos_q, q, None)
        q = ao.layer_norm(sd, "norm1.", q + q2)
        q = ao.layer_norm(sd, "norm_l.", q + ao.mha(sd, "cross_l.", q + pos_q, t, t, text_mask))
        q = ao.layer_norm(sd, "norm_d.", q + ao.mha(sd, "cross_d.", q + pos_q, det, det, det_mask))
        q = ao.layer_norm(sd, "norm_v.", q + ao.mha(sd, "cross_v.", q + pos_q, v, v, None))
        q = ao.layer_norm(sd, "norm2.", q + ao.ffn(sd, "ffn.", q))
    return q, v, t


def loss_of(out):
    q, v, t = out
    return q.pow(2).mean() + 0.1 * v.pow(2).mean() + 0.1 * t.pow(2).mean()


def timeit(fn, warmup, iters, world):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor([min(ts), sorted(ts)[len(ts) // 2]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"min_ms": t[0].item(), "med_ms": t[1].item()}


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, L, D, K = 8, 50000, 80, 132, 256
    g = torch.Generator().manual_seed(100 + rank)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    pc = synthetic.point_clouds(B, N, "surface", seed=synthetic.SEED + rank).to(dev)
    text, det, query, pos = r(B, L, ac.E), r(B, D, ac.E), r(B, K, ac.E), 0.5 * r(B, 1024, ac.E)
    text_mask = ac.ragged_mask(B, L, 20, g).to(dev)
    det_mask = ac.ragged_mask(B, D, 20, g).to(dev)
    qpos = torch.cat([4 * torch.rand(B, K, 3, generator=g) - 2, torch.rand(B, K, 3, generator=g) + .2], -1).to(dev)
    args = (pc, pos, text, text_mask, det, det_mask, query, qpos)

    torch.manual_seed(0)
    model = HotPath()
    ac.fill_params(model.encoder, 1)
    for i, d in enumerate(model.decoder):
        ac.fill_params(d, 10 + i)
    model = model.to(dev).train()
    ddp.broadcast_parameters(model)
    fg = ddp.FlatGradients(model)
    torch.backends.cuda.matmul.allow_tf32 = False

    def step():
        fg.zero()
        loss = loss_of(model(*args))
        loss.backward()
        fg.all_reduce_mean()
        return loss.detach()  # no reference to the autograd graph survives the step (CUDA-graph capture needs that)

    res = {"B_per_gpu": B, "world": world, "N": N, "gpu": torch.cuda.get_device_name(local),
           "trainable_params": sum(p.numel() for p in fg.params), "grad_bytes": fg.nbytes}
    loss = step()
    res["loss"] = loss.item()
    res["ours_fwd_bwd"] = timeit(step, 2, 5, world)
    res["ours_scenes_per_s"] = B * world / (res["ours_fwd_bwd"]["min_ms"] * 1e-3)
    with torch.no_grad():
        res["ours_fwd_only_train_mode"] = timeit(lambda: model(*args), 2, 5, world)
    eager_grads = fg.flat.clone()
    # the same step as ONE CUDA graph (eda_b200/graphs.py GraphedTrainStep); the all-reduce stays outside the graph
    try:
        from eda_b200.graphs import GraphedTrainStep

        gstep = GraphedTrainStep(model, loss_of, list(args), fg)

        def graphed():
            loss = gstep(*args)
            fg.all_reduce_mean()
            return loss

        res["graphed_loss"] = graphed().item()
        res["graphed_vs_eager_grad_rel_diff"] = ((fg.flat - eager_grads).norm() / eager_grads.norm()).item()
        res["ours_graphed_fwd_bwd"] = timeit(graphed, 3, 10, world)
        res["ours_graphed_scenes_per_s"] = B * world / (res["ours_graphed_fwd_bwd"]["min_ms"] * 1e-3)
    except Exception as e:  # noqa: BLE001
        res["graphed_error"] = repr(e)[:500]
        step()

    if world == 1:
        try:
            from oracle import ref_loader
            ref_ext = ref_loader.load_reference_ext()
        except Exception as e:  # noqa: BLE001
            ref_ext, res["ref_error"] = None, repr(e)
        if ref_ext is not None:
            ours_grads = {n: p.grad.clone() for n, p in model.named_parameters()}
            for sa in (model.backbone.sa1, model.backbone.sa2, model.backbone.sa3, model.backbone.sa4):
                sa.fuse = False
            model.backbone.overlap_fps = False

            def ref_step():
                fg.zero()
                loss = loss_of(reference_forward(model, ref_ext, *args))
                loss.backward()
                return loss.detach()

            res["ref_loss"] = ref_step().item()
            rel = {}
            for n, p in model.named_parameters():
                gr, go = p.grad, ours_grads[n]
                rel[n] = ((go - gr).norm() / gr.norm().clamp_min(1e-20)).item()
            worst = sorted(rel.items(), key=lambda kv: -kv[1])[:5]
            res["grad_rel_err_median"] = sorted(rel.values())[len(rel) // 2]
            res["grad_rel_err_worst5"] = worst
            res["ref_fwd_bwd"] = timeit(ref_step, 1, 3, 1)
            res["ref_scenes_per_s"] = B / (res["ref_fwd_bwd"]["min_ms"] * 1e-3)
            res["speedup_fwd_bwd"] = res["ref_fwd_bwd"]["min_ms"] / res["ours_fwd_bwd"]["min_ms"]
            if "ours_graphed_fwd_bwd" in res:
                res["speedup_fwd_bwd_graphed"] = res["ref_fwd_bwd"]["min_ms"] / res["ours_graphed_fwd_bwd"]["min_ms"]
    if rank == 0:
        print(json.dumps(res, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
