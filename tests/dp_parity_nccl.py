"""torchrun-able NCCL data-parallel parity check: W ranks on row shards + bucketed all-reduce must give
the gradients one rank computes on the same global batch (rtol 1e-5 / atol 1e-6 of the tensor scale).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_parity_nccl.py

Launched by tests/test_config_parity.py::test_nccl_data_parallel_gradients_equal_single_rank when >= 2 GPUs exist;
`bench.py --gpus N` runs the same function and reports it as `dp_parity`."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from auto_differentiation_b200.data_parallel import dp_parity
    from auto_differentiation_b200.models.model import AttentionClassifier
    rng = np.random.default_rng(77)
    B, S, d = 22, 256, 512                                  # uneven shards (22 rows over W ranks), fused d_k = 64 attention
    x = rng.standard_normal((B, S, d)).astype(np.float32)
    t = rng.integers(0, 10, B)
    res = dp_parity(lambda: AttentionClassifier(d_model=d, heads=8, hidden=256, classes=10, micro_batch=4, seed=5), x, t)
    if dist.get_rank() == 0:
        print(json.dumps(res))
        assert res["ok"], res
        print("dp_parity ok")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
