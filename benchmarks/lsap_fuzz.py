"""Fuzzes the device LSAP kernel against scipy on tie-heavy integer matrices and random float matrices, batched and
ragged (the GPU tests cover 39 fixed matrices; this runs thousands).  Prints one JSON line with the mismatch count."""
import json
import os
import sys

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mp_former_b200 import native  # noqa: E402


def main(batches=int(os.environ.get("MPF_FUZZ_BATCHES", "60")), per=32):
    rng = np.random.default_rng(7)
    bad, total = 0, 0
    for it in range(batches):
        Q = int(rng.integers(1, 130))
        counts = [int(rng.integers(0, 140)) for _ in range(per)]
        kind = it % 3
        mats = [(rng.integers(0, int(rng.integers(1, 4)) + 1, (Q, n)).astype(np.float32) if kind < 2
                 else rng.standard_normal((Q, n)).astype(np.float32)) for n in counts]
        flat = torch.from_numpy(np.concatenate([m.reshape(-1) for m in mats])).cuda()
        offs = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32, device="cuda")
        qi, ti, status = native.lsap(flat, offs, counts, Q)
        assert int(status.item()) == 0
        sizes = [min(Q, n) for n in counts]
        for i, j, M in zip(torch.split(qi.cpu(), sizes), torch.split(ti.cpu(), sizes), mats):
            total += 1
            if M.shape[1] == 0:
                continue
            ri, ci = linear_sum_assignment(M)
            if not (np.array_equal(i.numpy(), ri) and np.array_equal(j.numpy(), ci)):
                bad += 1
    print(json.dumps({"probe": "lsap_fuzz_vs_scipy", "matrices": total, "mismatches": bad}))


if __name__ == "__main__":
    main()
