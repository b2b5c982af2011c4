#!/usr/bin/env python
"""Generates the trained-weight fixtures of the reference's CAST checkpoints (build container only; reads
/root/reference/saved_models/ml-1m.txt/cast_{1..6}_*/model.ckpt with this repo's TF-free bundle reader):

  ml1m_cast{1..6}_ckpt.npz   variables under role names (checkpoint.name_map: mapped by role; the context tower's
                             dead `ln` pair is checked to be still at its initialiser and dropped) + global_step
  ref_bundle_crc.npz         for the SASRec checkpoint: the masked crc32c field TensorFlow stored in model.ckpt.index
                             for every tensor (known answers for the writer's checksum, byte level)

    python tests/golden/make_cast_ckpt_golden.py
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cast_b200  # noqa: E402,F401
from cast_b200 import checkpoint as ck  # noqa: E402

ROOT = "/root/reference/saved_models/ml-1m.txt"


def main():
    for n in range(1, 7):
        d = glob.glob(os.path.join(ROOT, f"cast_{n}_*"))[0]
        tfv = ck.read_bundle(os.path.join(d, "model.ckpt"))
        nm = ck.name_map(f"cast_{n}", 2)
        for tf_name, role in nm.items():
            if role.startswith(ck.DEAD):   # never trained: beta == 0, gamma == 1
                want = 1.0 if role.endswith("gamma") else 0.0
                assert np.all(tfv[tf_name] == want), tf_name
        roles = ck.to_role_names(f"cast_{n}", tfv, 2)
        np.savez_compressed(os.path.join(HERE, f"ml1m_cast{n}_ckpt.npz"), global_step=tfv["global_step"], **roles)
        print(f"cast_{n}: {len(roles)} tensors, global_step {int(tfv['global_step'])}")
    prefix = os.path.join(ROOT, "sasrec_baseline_10-19-2019-21-23-42", "model.ckpt")
    crcs = ck.read_bundle_crcs(prefix)
    np.savez_compressed(os.path.join(HERE, "ref_bundle_crc.npz"), names=np.array(sorted(crcs)),
                        crcs=np.array([crcs[k] for k in sorted(crcs)], dtype=np.uint32))
    print("sasrec index: masked crc32c of", len(crcs), "tensors")


if __name__ == "__main__":
    main()
