"""torchrun sanity check of the data-parallel step: ranks train on DIFFERENT samples; after every step all ranks
must hold bit-identical parameters, and the bucketed (head under towers' backward) all-reduce must give the same
parameters as eager forward/backward + one all-reduce.  usage: torchrun --nproc-per-node 2 tools/dp_check.py"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monopsr_b200.core import model_spec as ms
from monopsr_b200.core.engine import Engine

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
P = ms.init_params(0)
S = ms.synthetic_sample(rank)                 # a different sample per rank
a, b = Engine(dev, params=P), Engine(dev, params=P)
a.set_inputs(S); b.set_inputs(S)
ok = True
for step in range(3):
    a.train_step()                            # default path: four graphs around two all-reduces
    b.set_hyper(b.step_count); b.train_step_eager(); b.step_count += 1
    torch.cuda.synchronize()
    pa, pb = a.params, b.params
    same = [torch.zeros_like(pa) for _ in range(world)]
    dist.all_gather(same, pa)
    ident = all(bool(torch.equal(same[0], t)) for t in same)
    rel = float((pa - pb).norm() / pb.norm())
    if rank == 0:
        print("step %d: ranks identical %s; graph/bucketed vs eager/one all-reduce rel diff %.3e; finite %s" % (
            step, ident, rel, bool(torch.isfinite(pa).all())))
    ok = ok and ident and rel < 1e-3      # graph vs eager differ by the order of fp32 RED.ADDs, amplified by Adam
# timing of the three variants on this box (device events, max over ranks)
def timed(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
t_split = timed(a.train_step)
t_graph = float("nan")
if os.environ.get("MPB_DP_CHECK_GRAPH", "0") == "1":
    os.environ["MPB_DP_GRAPH"] = "1"
    c = Engine(dev, params=P); c.set_inputs(S)
    t_graph = timed(c.train_step)
if rank == 0:
    print("ms/step at %d ranks: four graphs, head opt under the tower all-reduce %.3f | one graph with bucketed all-reduces %.3f" % (world, t_split, t_graph))
    print("DP CHECK", "OK" if ok else "FAILED", flush=True)
sys.stdout.flush()
torch.cuda.synchronize()
os._exit(0)        # (communicator teardown with live CUDA graphs can block for minutes)
