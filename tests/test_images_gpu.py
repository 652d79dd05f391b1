"""GPU parity tests of the image path: sx_sextans_invoke fed with the reference's own
FPGA channel images (oracle/_ref) or the committed golden images, read back the way the
host program reads its result (src/sextans-host.cpp:264-270), against cpu_spmm_CSR --
bit for bit -- and the UNMODIFIED reference host program running on the engine."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import GOLDEN, mtx_path, random_csr, random_dense, sha

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = oracle.ref() is not None and hasattr(oracle.ref(), "sxref_build_images")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libsextans_ref.so not built")


@pytest.fixture(scope="module")
def eng():
    e = sx.Engine(0)
    yield e
    e.close()


def invoke(eng, ptr, imgs, M, K, N, alpha, beta, B, Cin, rp_time=1):
    P_N, au, bu = oracle.pack_scalars(N, rp_time, alpha, beta)
    bi = oracle.pack_B_images(B, K, N)
    ci = oracle.pack_C_images(Cin, M, N)
    co = [np.zeros_like(x) for x in ci]
    ns = eng.sextans_invoke(ptr, imgs, bi, ci, co, M, K, P_N, au, bu)
    return oracle.unpack_C_images(co, M, N), ns


def test_golden_images_through_the_engine(eng, golden):
    g = np.load(os.path.join(GOLDEN, "images_small.npz"))
    for tag in golden["image_cases"]:
        M, K, _ = g[tag + "_dims"].tolist()
        N = 16
        B, Cin = random_dense(M, K, N, 31, np.float32)
        C, ns = invoke(eng, g[tag + "_ptr"], [g[f"{tag}_A{c}"] for c in range(8)], M, K, N, 0.85, -2.06, B, Cin)
        ref = oracle.spmm_csr(M, N, K, g[tag + "_rowptr"], g[tag + "_colidx"], g[tag + "_val"],
                              np.float32(0.85), B, np.float32(-2.06), Cin.copy())
        assert np.array_equal(C.view(np.uint32), ref.view(np.uint32)), tag
        assert ns > 0


@needs_ref
def test_nasa4704_canned_run_from_images(eng, golden):
    """The reference's canned run (CMakeLists.txt:49: nasa4704, N=16, defaults) with the
    reference's own images: the result read back from the C images hashes to the golden
    cpu_spmm_CSR output."""
    M, K, nnz, rp, ci, v = oracle.ref_load_csr(mtx_path("nasa4704"))
    ptr, imgs, num_a_len = oracle.ref_build_images(M, K, rp, ci, v)
    B, Cin = oracle.init_dense(M, K, 16, np.float32)
    serial0 = eng.info(sx.INFO_UPLOAD_SERIAL)
    C, ns = invoke(eng, ptr, imgs, M, K, 16, 0.85, -2.06, B, Cin)
    run = [r for r in golden["suitesparse"]["nasa4704"]["runs"] if r["kind"] == "default" and r["N"] == 16][0]
    assert sha(C) == run["C_sha256"]
    serial1 = eng.info(sx.INFO_UPLOAD_SERIAL)
    assert serial1 != serial0 and eng.info(sx.INFO_NNZ) == nnz
    # same images again (rp_time = 3): A is not decoded/uploaded again, same result
    C2, ns3 = invoke(eng, ptr, imgs, M, K, 16, 0.85, -2.06, B, Cin, rp_time=3)
    assert eng.info(sx.INFO_UPLOAD_SERIAL) == serial1
    assert np.array_equal(C2.view(np.uint32), C.view(np.uint32)) and ns3 > 0
    # other values in the same pattern: detected, uploaded, and the answer follows
    v2 = (1.0 + 0.001 * (np.arange(nnz) % 97)).astype(np.float32)
    ptr2, imgs2, _ = oracle.ref_build_images(M, K, rp, ci, v2)
    C3, _ = invoke(eng, ptr2, imgs2, M, K, 16, 0.85, -2.06, B, Cin)
    assert eng.info(sx.INFO_UPLOAD_SERIAL) != serial1
    ref = oracle.spmm_csr(M, 16, K, rp, ci, v2, np.float32(0.85), B, np.float32(-2.06), Cin.copy())
    assert np.array_equal(C3.view(np.uint32), ref.view(np.uint32))


@needs_ref
@pytest.mark.parametrize("M,K,avg,N,alpha,beta", [(300, 9000, 12, 24, 0.85, -2.06), (1, 5, 3, 8, 1.0, 0.0),
                                                  (1000, 4097, 3, 8, -1.25, 0.5), (129, 700, 9, 64, 0.0, 1.5),
                                                  (50, 40, 0, 16, 0.85, -2.06)])
def test_random_images_bit_exact(eng, M, K, avg, N, alpha, beta):
    rp, ci, v = random_csr(M, K, avg, M + K, np.float32)
    ptr, imgs, _ = oracle.ref_build_images(M, K, rp, ci, v)
    B, Cin = random_dense(M, K, N, 7, np.float32)
    C, _ = invoke(eng, ptr, imgs, M, K, N, alpha, beta, B, Cin)
    ref = oracle.spmm_csr(M, N, K, rp, ci, v, np.float32(alpha), B, np.float32(beta), Cin.copy())
    assert np.array_equal(C.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("tag", ["p", "q", "r"])
def test_output_images_equal_the_dataflow_model_word_for_word(eng, tag):
    """Not only the rows the host reads back: every float of the output images the hardware
    would write (whole 16-row words, pad rows included) equals the functional model of the
    accelerator dataflow run on the same images (oracle.sextans_images)."""
    g = np.load(os.path.join(GOLDEN, "images_small.npz"))
    M, K, _ = g[tag + "_dims"].tolist()
    ptr, imgs = g[tag + "_ptr"], [g[f"{tag}_A{c}"] for c in range(8)]
    for N, alpha, beta, rp_time in ((8, 0.85, -2.06, 1), (24, -1.25, 0.5, 3)):
        B, Cin = random_dense(M, K, N, 23, np.float32)
        P_N, au, bu = oracle.pack_scalars(N, rp_time, alpha, beta)
        bi = oracle.pack_B_images(B, K, N)
        ci = oracle.pack_C_images(Cin, M, N, fill_pad=0.25)
        co = [np.full_like(x, 7.0) for x in ci]
        eng.sextans_invoke(ptr, imgs, bi, ci, co, M, K, P_N, au, bu)
        model = oracle.sextans_images(ptr, imgs, bi, ci, M, K, P_N, au, bu)
        used = sx.lib().sx_images_C_floats(M, N)
        for o, e in zip(co, model):
            assert np.array_equal(o[:used].view(np.uint32), e[:used].view(np.uint32))
            assert (o[used:] == 7.0).all()


def test_user_upload_between_invokes_is_noticed(eng, golden):
    g = np.load(os.path.join(GOLDEN, "images_small.npz"))
    M, K, _ = g["q_dims"].tolist()
    ptr, imgs = g["q_ptr"], [g[f"q_A{c}"] for c in range(8)]
    B, Cin = random_dense(M, K, 8, 3, np.float32)
    C1, _ = invoke(eng, ptr, imgs, M, K, 8, 0.5, 2.0, B, Cin)
    rp, ci, v = random_csr(M, K, 3, 99, np.float32)
    eng.upload_csr(M, K, rp, ci, v)  # same shape, other matrix, behind the image path's back
    C2, _ = invoke(eng, ptr, imgs, M, K, 8, 0.5, 2.0, B, Cin)
    assert np.array_equal(C1.view(np.uint32), C2.view(np.uint32))


def test_the_unmodified_reference_host_program_runs_on_the_engine():
    """src/sextans-host.cpp as it is (built into oracle/_ref/sextans_ref_host with
    include/tapa_compat and integration/sextans_kernel_b200.cpp): its own loader, FPGA
    preprocessing, channel repacking and verification around our device call."""
    exe = os.path.join(ROOT, "oracle", "_ref", "sextans_ref_host")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/sextans_ref_host not built")
    for args in (["16"], ["8", "4"], ["24", "2", "1.5", "0.25"]):
        r = subprocess.run([exe, mtx_path("nasa4704")] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert "Success!" in r.stdout
        m = re.search(r"num_mismatch = (\d+)", r.stdout)
        assert m and int(m.group(1)) == 0, r.stdout[-400:]
        assert re.search(r"Kernel time is [0-9.]+ ms", r.stdout)
