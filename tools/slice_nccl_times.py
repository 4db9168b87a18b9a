"""One 2^lg-term MSM over all ranks (torchrun): point-range shards (ShardedMsm) against bucket slices over a replicated table
(SlicedMsm, one- and two-part forms), every result checked against the closed form.  Development aid behind
profiles/r02c_bucket_slices_and_scan_reduce.md."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import porla_b200 as pb
import bench
from oracle import curves_py as O
from porla_b200.sharding import ShardedMsm, SlicedMsm, shard_range

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
lib = pb.load(); lib.porla_device_init()


def timed(fn, reps=8):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for lg in [int(x) for x in sys.argv[1:]] or [24]:
    n = 1 << lg
    parts = [bench.strong_shard_inputs(torch, lg, world, r, dev) for r in range(world)]
    ks_all = torch.cat([p[2] for p in parts]); ss_all = torch.cat([p[3] for p in parts])
    total = sum(bench.weighted_scalar_sum(torch, p[3], p[0]) for p in parts) % O.BN254.n
    del parts
    want = O.bn254_marshal(O.mul(O.BN254, total, (1, 2)))
    lo, hi = shard_range(n, world, rank)
    own = ss_all[lo:hi].contiguous()
    res = [None]
    tab_r = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks_all[lo:hi].contiguous().data_ptr(), hi - lo, pb.SCALAR_LE32, on_device=True)
    eng = ShardedMsm(pb.CURVE_BN254, n, world, rank, dist, dev)
    def f_range(): res[0] = eng.msm(tab_r, own.data_ptr(), hi - lo, pb.SCALAR_LE32)
    ms_range = timed(f_range); ok_range = rank != 0 or res[0] == want
    tab_r.destroy()
    tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks_all.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
    del ks_all, ss_all
    out = {}
    for two in (False, True):
        e2 = SlicedMsm(pb.CURVE_BN254, n, world, rank, dist, dev, two_part=two)
        def f_slice(): res[0] = e2.msm(tab, own, pb.SCALAR_LE32)
        out[two] = (timed(f_slice), rank != 0 or res[0] == want)
        del e2
    tab.destroy()
    if rank == 0:
        print("2^%d over %d GPUs: range shards %.3f ms (%s) | bucket slices, one part %.3f ms (%s), own range first %.3f ms (%s)" %
              (lg, world, ms_range, ok_range, out[False][0], out[False][1], out[True][0], out[True][1]), flush=True)
dist.barrier()
dist.destroy_process_group()
