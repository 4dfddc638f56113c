"""World-size-2 gloo run (CPU) of the multi-GPU plumbing of bench.py: image sharding by rank, max-over-ranks
timing, whole-job aggregation.  The data path itself has no collective."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %(root)r)
    import torch, torch.distributed as dist
    import bench
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # rank r owns images [16 r, 16 r + 16): the alpha recipe is seeded per global image index
    a = bench.synthetic_alpha(2, seed0=bench.BATCH_PER_GPU * rank)
    ids = torch.tensor([bench.BATCH_PER_GPU * rank + i for i in range(2)])
    gathered = [torch.zeros_like(ids) for _ in range(world)]
    dist.all_gather(gathered, ids)
    # per-rank step time -> MAX over ranks -> whole-job images/s
    t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = bench.BATCH_PER_GPU * world / (float(t) * 1e-3)
    chk = torch.tensor([float(a.sum())], dtype=torch.float64)
    sums = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(sums, chk)
    if rank == 0:
        print(json.dumps({"ids": torch.cat(gathered).tolist(), "ms": float(t), "value": value,
                          "alpha_sums": [float(s) for s in sums]}))
    dist.destroy_process_group()
""")


def test_image_sharding_and_aggregation_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["ids"] == [0, 1, 16, 17]                      # disjoint image shards
    assert d["ms"] == 15.0                                 # max over ranks
    assert abs(d["value"] - 32 / 15e-3) < 1e-6             # whole-job images/s
    assert d["alpha_sums"][0] != d["alpha_sums"][1]        # different images on different ranks


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


DP_WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %(root)r)
    import torch, torch.distributed as dist
    import mwa_b200
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(0)                                   # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(12, 40), torch.nn.Tanh(), torch.nn.Linear(40, 3))
    frozen = torch.nn.Parameter(torch.ones(5), requires_grad=False)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(8, 12, generator=g) * 3
    Y = torch.randn(8, 3, generator=g) * 40                # large targets -> some gradients beyond the +-5 clip
    # reference: one process, whole batch, clip after backward (trainRGB.py:186-196)
    ref = torch.nn.Sequential(torch.nn.Linear(12, 40), torch.nn.Tanh(), torch.nn.Linear(40, 3))
    ref.load_state_dict(model.state_dict())
    torch.nn.functional.mse_loss(ref(X), Y).backward()
    for p in ref.parameters():
        p.grad.clamp_(-5, 5)
    # data parallel: rank r owns samples [4 r, 4 r + 4)
    xs, ys = X[4 * rank:4 * rank + 4], Y[4 * rank:4 * rank + 4]
    torch.nn.functional.mse_loss(model(xs), ys).backward()
    ar = mwa_b200.GradientAllReduce(list(model.parameters()) + [frozen], bucket_bytes=1024, clip_value=5.0)
    ar()
    err = max(float((p.grad - q.grad).abs().max()) for p, q in zip(model.parameters(), ref.parameters()))
    clipped = sum(int((q.grad.abs() == 5).sum()) for q in ref.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=1e-2); opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    if rank == 0:
        print(json.dumps({"err": err, "clipped": clipped, "buckets": len(ar.buckets), "bytes": ar.bytes_per_step,
                          "replicas_equal": bool(torch.equal(both[0], both[1]))}))
    dist.destroy_process_group()
""")


def test_gradient_allreduce_matches_single_process_world2(tmp_path):
    """config 5's only collective: bucketed gradient all-reduce + clip AFTER the reduction == the reference's
    single-process step on the global batch; replicas stay bit-identical after the optimizer step."""
    script = tmp_path / "dp_worker.py"
    script.write_text(DP_WORKER % dict(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["err"] < 1e-5, d
    assert d["clipped"] > 0, "the recipe must exercise the clip"
    assert d["buckets"] > 1 and d["bytes"] == 4 * (12 * 40 + 40 + 40 * 3 + 3)
    assert d["replicas_equal"]
