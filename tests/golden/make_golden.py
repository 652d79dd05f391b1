#!/usr/bin/env python
"""Generate the golden fixtures from the REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference and the compiled
oracle/_ref/libsextans_ref.so = the reference's src/sparse_helper.h + src/mmio.h,
unmodified).  Nothing here calls the oracle restatement: every number written
comes out of the reference's read_suitsparse_matrix / CSC_2_CSR / cpu_spmm_CSR.

Outputs (committed):
  matrices/<name>.mtx.xz   the two SuiteSparse sample matrices the reference ships
                           (UF Sparse Matrix Collection, Boeing/nasa4704 and
                           Boeing/pcrystk02) -- data, xz-compressed
  loader_small.npz         reference CSR (rowptr/colidx/val) of every tests/golden/mtx/*.mtx
  golden.json              sizes, row-length stats, sha256 of the reference CSR and of
                           the reference C for the SuiteSparse runs, known answers
  spmm_small.npz           full reference C for small seeded cases, with their inputs
  images_small.npz         the reference's FPGA channel images (generate_edge_list_for_all_PEs
                           + edge_list_64bit, src/sparse_helper.h:345-473) of small seeded
                           matrices, with the CSR they were made from
                           (`make_golden.py --images-only` rewrites just this file)
"""
import json
import lzma
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from helpers import SMALL_MTX, SUITESPARSE, perturbed_inputs, random_csr, random_dense, sha  # noqa: E402

REF = "/root/reference"


IMAGE_SPECS = [  # (tag, M, K, avg nnz/row, long_row)
    ("p", 70, 5000, 4, None),     # two column windows
    ("q", 130, 64, 7, 60),        # more rows than PEs, one dense row
    ("r", 3, 9000, 2, None),      # fewer rows than PEs, three windows
    ("s", 65, 4096, 0, None),     # no nonzeros at all
]


def write_images():
    out = {}
    for i, (tag, M, K, avg, long_row) in enumerate(IMAGE_SPECS):
        rp, ci, v = random_csr(M, K, avg, 200 + i, np.float32, long_row=long_row)
        ptr, imgs, num_a_len = oracle.ref_build_images(M, K, rp, ci, v)
        out[tag + "_dims"] = np.array([M, K, num_a_len], dtype=np.int64)
        out[tag + "_ptr"] = ptr
        for c in range(8):
            out[f"{tag}_A{c}"] = imgs[c]
        for k, a in (("rowptr", rp), ("colidx", ci), ("val", v)):
            out[f"{tag}_{k}"] = a
    np.savez_compressed(os.path.join(HERE, "images_small.npz"), **out)


def main():
    assert oracle.ref() is not None, "build oracle/_ref first (make -C oracle)"
    if "--images-only" in sys.argv:
        write_images()
        return
    gold = {"generator": "tests/golden/make_golden.py", "source": "reference cpu_spmm_CSR "
            "(src/sparse_helper.h:262-290) via oracle/_ref/libsextans_ref.so", "suitesparse": {}}

    os.makedirs(os.path.join(HERE, "matrices"), exist_ok=True)
    for name in SUITESPARSE:
        src = f"{REF}/matrices/{name}/{name}.mtx"
        with open(src, "rb") as f:
            raw = f.read()
        with open(os.path.join(HERE, "matrices", name + ".mtx.xz"), "wb") as f:
            f.write(lzma.compress(raw, preset=9 | lzma.PRESET_EXTREME))
        M, K, nnz, rp, ci, v = oracle.ref_load_csr(src)
        lens = np.diff(rp)
        entry = {"M": M, "K": K, "nnz": nnz, "mtx_sha256": sha(np.frombuffer(raw, np.uint8)),
                 "rowptr_sha256": sha(rp), "colidx_sha256": sha(ci), "val_sha256": sha(v),
                 "row_len_min": int(lens.min()), "row_len_max": int(lens.max()), "runs": []}
        # (1) the canned run of the reference: alpha=0.85 beta=-2.06, B=1, C=(m+1)(n+1)/M/N
        for N in (8, 16, 32, 64):
            B, C = oracle.init_dense(M, K, N, np.float32)
            oracle.ref_spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, C)
            entry["runs"].append({"kind": "default", "N": N, "alpha": 0.85, "beta": -2.06,
                                  "C0": float(C[0]), "C_Mm1": float(C[M - 1]),
                                  "C_last": float(C[-1]),
                                  "sum": float(C.astype(np.float64).sum()),
                                  "C_sha256": sha(C)})
        # (2) value-sensitive run (pattern values overwritten)
        for N in (8, 16):
            val, B, C = perturbed_inputs(M, K, N, nnz, np.float32)
            oracle.ref_spmm_csr(M, N, K, rp, ci, val, 0.85, B, -2.06, C)
            entry["runs"].append({"kind": "perturbed", "N": N, "alpha": 0.85, "beta": -2.06,
                                  "C0": float(C[0]), "C_Mm1": float(C[M - 1]),
                                  "C_last": float(C[-1]),
                                  "sum": float(C.astype(np.float64).sum()),
                                  "C_sha256": sha(C)})
        # the FPGA scheduler's padded slot count (2140 / 15489)
        entry["edge_list_slots"] = int(oracle.ref().sxref_edge_list_slots(
            M, K, nnz, rp.ctypes.data_as(oracle._PI), ci.ctypes.data_as(oracle._PI),
            v.ctypes.data_as(oracle._PF), 64, 4096, 10))
        gold["suitesparse"][name] = entry

    small = {}
    for name in SMALL_MTX:
        M, K, nnz, rp, ci, v = oracle.ref_load_csr(os.path.join(HERE, "mtx", name + ".mtx"))
        small[name + "_shape"] = np.array([M, K, nnz], dtype=np.int64)
        small[name + "_rowptr"] = rp
        small[name + "_colidx"] = ci
        small[name + "_val"] = v
    np.savez_compressed(os.path.join(HERE, "loader_small.npz"), **small)

    cases = {}
    specs = [  # (tag, M, K, avg, N, alpha, beta, long_row)
        ("a", 64, 48, 6, 8, 0.85, -2.06, None),
        ("b", 33, 70, 3, 16, 1.0, 0.0, 60),
        ("c", 17, 9, 4, 24, 0.0, 1.5, None),
        ("d", 128, 128, 20, 32, -1.25, 0.5, None),
        ("e", 1, 5, 3, 8, 2.0, 2.0, None),
        ("f", 50, 40, 0, 8, 0.85, -2.06, None),  # all rows empty
    ]
    for i, (tag, M, K, avg, N, alpha, beta, long_row) in enumerate(specs):
        rp, ci, v = random_csr(M, K, avg, 100 + i, np.float32, long_row=long_row)
        B, Cin = random_dense(M, K, N, 100 + i, np.float32)
        C = Cin.copy()
        oracle.ref_spmm_csr(M, N, K, rp, ci, v, alpha, B, beta, C)
        cases[tag + "_dims"] = np.array([M, K, N], dtype=np.int64)
        cases[tag + "_ab"] = np.array([alpha, beta], dtype=np.float32)
        for k, a in (("rowptr", rp), ("colidx", ci), ("val", v), ("B", B), ("Cin", Cin), ("C", C)):
            cases[f"{tag}_{k}"] = a
    np.savez_compressed(os.path.join(HERE, "spmm_small.npz"), **cases)
    gold["small_cases"] = [s[0] for s in specs]
    write_images()
    gold["image_cases"] = [s[0] for s in IMAGE_SPECS]

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote golden fixtures:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
