"""Worker of tests/test_engine_gpu.py::test_two_gpu_shard_parity (launched with torch.distributed.run, one rank per GPU).

SURVEY.md section 8(e): rank r's outputs equal a single-GPU run on shard r (BatchNorm statistics are local), and the
reduced gradient equals the mean of the per-shard single-GPU gradients within 1e-3; then the graphed step (NCCL
all-reduce captured inside the CUDA graph) must reproduce the eager distributed step."""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from scan2cap_b200 import synthetic
    from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig
    from scan2cap_b200.distributed import shard_batch
    from scan2cap_b200.engine import TrainStep
    from scan2cap_b200.models.capnet import CapNet
    DC = ScannetDatasetConfig()
    V, per = 200, 2
    vocab, emb, _ = synthetic.make_vocabulary(V)
    cfg = dict(input_feature_dim=4, num_proposal=256, num_locals=10, use_topdown=True, query_mode="center",
               graph_mode="edge_conv", num_graph_steps=2, use_relation=True, use_orientation=True)
    torch.manual_seed(0)
    base = CapNet(DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr, **cfg).to(dev)
    base.train()
    d = synthetic.make_data_dict(per * world, 8000, use_normal=True, num_vocabs=V, seed=21)
    full = {k: torch.from_numpy(v).to(dev) for k, v in d.items()}
    with torch.no_grad():
        probe = copy.deepcopy(base)(dict(full))
    full["ref_box_corner_label"] = probe["bbox_corner"][:, 5].clone()
    num_words = int(d["lang_len"].max())
    flags = dict(detection=True, caption=True, orientation=True, distance=False)
    res = {}

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-12))

    # --- single-GPU runs on every shard (no collective), same initial weights
    shard_grads, shard_loss = [], []
    for r in range(world):
        m = copy.deepcopy(base)
        eng = TrainStep(m, DC, use_cuda_graph=False, **flags)
        eng.world = 1
        out = eng._fwd_bwd(dict(shard_batch(full, r, world), num_words=num_words))
        shard_grads.append(eng.flat.flat.clone())
        shard_loss.append(out["loss"].detach().clone())
        if r == rank:
            mine_single = {k: out[k].detach().clone() for k in ("sa1_inds", "aggregated_vote_inds", "lang_cap", "loss")}
    want = sum(shard_grads) / world

    # --- distributed eager step on this rank's shard
    m = copy.deepcopy(base)
    eng = TrainStep(m, DC, use_cuda_graph=False, **flags)
    out = eng._fwd_bwd(dict(shard_batch(full, rank, world), num_words=num_words))
    res["rank_output_equals_single_gpu_shard"] = bool(
        torch.equal(out["sa1_inds"], mine_single["sa1_inds"]) and
        torch.equal(out["aggregated_vote_inds"], mine_single["aggregated_vote_inds"]) and
        rel(out["lang_cap"], mine_single["lang_cap"]) < 1e-5 and rel(out["loss"], mine_single["loss"]) < 1e-5)
    eng.flat.all_reduce_mean()
    res["reduced_grad_vs_mean_of_shards"] = float((eng.flat.flat.double() - want.double()).norm() / want.double().norm())
    eng.opt.step()
    eager_params = torch.cat([p.detach().flatten() for p in m.parameters()])

    # --- graphed distributed step (all-reduce captured inside the graph): same update as the eager one
    m2 = copy.deepcopy(base)
    eng2 = TrainStep(m2, DC, use_cuda_graph=True, word_bucket=1, **flags)
    loss2 = eng2.run(dict(shard_batch(full, rank, world), num_words=num_words))
    torch.cuda.synchronize()
    res["graph_loss_vs_eager"] = rel(loss2, out["loss"].detach())
    graph_params = torch.cat([p.detach().flatten() for p in m2.parameters()])
    moved = (eager_params - torch.cat([p.detach().flatten() for p in base.parameters()])).abs() > 5e-4
    res["graph_update_sign_agreement"] = float(((graph_params - eager_params).abs()[moved] < 2e-4).float().mean())
    res["graph_grad_vs_mean_of_shards"] = float((eng2.flat.flat.double() - want.double()).norm() / want.double().norm())
    steps = [int(s["step"]) for s in eng2.opt.state.values()]
    res["adam_steps_after_first_graph_run"] = [min(steps), max(steps)]
    res["bn_batches_tracked"] = int(m2.backbone_net.sa1.mlp_module.layer0.bn.bn.num_batches_tracked)
    # second step: replay only
    loss3 = eng2.run(dict(shard_batch(full, rank, world), num_words=num_words))
    torch.cuda.synchronize()
    res["second_step_loss_finite"] = bool(torch.isfinite(loss3))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        print("RESULT " + json.dumps(gathered), flush=True)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)   # destroy_process_group() blocks while captured graphs hold the communicator's kernels


if __name__ == "__main__":
    main()
