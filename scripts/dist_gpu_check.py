"""torchrun check of the item-sharded CUDA path (NCCL): ShardedB200Ranker over WORLD_SIZE GPUs == fp64 oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.topk_oracle import rank_oracle  # noqa: E402
from rectools_b200.sharded import ShardedB200Ranker  # noqa: E402
from tests.helpers import synth_factors, synth_viewed_csr  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_users, n_items, d, k = 3000, 41_003, 64, 10
    u, i = synth_factors(n_users, n_items, d, seed=21)
    csr = synth_viewed_csr(n_users, n_items, 40)
    whitelist = np.sort(np.random.default_rng(3).choice(n_items, 20_000, replace=False))
    ok = True
    for dist_name in ("dot", "cosine"):
        ranker = ShardedB200Ranker(dist_name, u, i)
        for filt, wl in ((csr, None), (csr, whitelist), (None, None)):
            sids = np.arange(n_users)
            s, ids, sc = ranker.rank(sids, k, filt, wl)
            if dist.get_rank() == 0:
                sel = sids[::7]
                es, eid, esc = rank_oracle(dist_name, u, i, sel, k, None if filt is None else filt[sel], wl, accum="f64")
                got = np.isin(s, sel)
                same = np.array_equal(ids[got], eid) and np.allclose(sc[got], esc, rtol=1e-6, atol=1e-7)
                print(f"{dist_name} filter={filt is not None} whitelist={wl is not None}: {'OK' if same else 'MISMATCH'}", flush=True)
                ok = ok and same
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
