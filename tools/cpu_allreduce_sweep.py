#!/usr/bin/env python
"""CPU allreduce latency: our TCP-mesh CPU backend vs torch gloo (world 2 and 4), wall clock, fp32 SUM.  python tools/cpu_allreduce_sweep.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
def worker(rank, world, port):
    import torch, torch.distributed as td
    import pytorch_distributed_train_b200 as pdt
    torch.set_num_threads(2)
    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    out=[]
    for n in [256, 29034, 1<<18, 1<<20, 1<<22, 1<<24]:
        t=torch.ones(n)
        def tm(fn, it):
            for _ in range(3): fn()
            td.barrier(); t0=time.perf_counter()
            for _ in range(it): fn()
            td.barrier(); return (time.perf_counter()-t0)/it*1e3
        it = 200 if n < 1<<20 else 20
        a=tm(lambda: pdt.distributed.all_reduce(t), it)
        b=tm(lambda: td.all_reduce(t), it)
        out.append((n*4, round(a,3), round(b,3)))
    td.destroy_process_group()
    return out
if __name__ == '__main__':
    from mp_helpers import run_ranks, free_port
    for w in (2,4):
        r=run_ranks(worker, w, free_port())[0]
        print('world',w, 'bytes, ours ms, gloo ms:', r)
