"""CPU tests of the image path's host side (no GPU): the decoder of the FPGA channel
images against the reference's own encoder (oracle/_ref, when present) and against the
committed golden images, the B/C image layouts against the restated host-program loops
(oracle.pack_* cite src/sextans-host.cpp), and the error behaviour."""
import os
import subprocess

import numpy as np
import pytest

import oracle
import sextans_b200 as sx
from helpers import GOLDEN, SUITESPARSE, mtx_path, random_csr, random_dense

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = oracle.ref() is not None and hasattr(oracle.ref(), "sxref_build_images")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libsextans_ref.so not built")


def same_csr(a, b):
    return (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32)))


def test_golden_images_decode_to_the_csr_they_were_made_from(golden):
    g = np.load(os.path.join(GOLDEN, "images_small.npz"))
    for tag in golden["image_cases"]:
        M, K, num_a_len = g[tag + "_dims"].tolist()
        ptr = g[tag + "_ptr"]
        assert ptr[-1] == num_a_len and ptr.size == (K + 4095) // 4096 + 1
        imgs = [g[f"{tag}_A{c}"] for c in range(8)]
        got = sx.images_decode_A(ptr, imgs, M, K)
        assert same_csr(got, (g[tag + "_rowptr"], g[tag + "_colidx"], g[tag + "_val"])), tag


@needs_ref
@pytest.mark.parametrize("name", SUITESPARSE)
def test_suitesparse_images_decode_to_the_reference_csr(name, golden):
    M, K, nnz, rp, ci, v = oracle.ref_load_csr(mtx_path(name))
    v = (1.0 + 0.001 * (np.arange(nnz) % 97)).astype(np.float32)  # value-sensitive
    ptr, imgs, num_a_len = oracle.ref_build_images(M, K, rp, ci, v)
    assert num_a_len == golden["suitesparse"][name]["edge_list_slots"]
    assert same_csr(sx.images_decode_A(ptr, imgs, M, K), (rp, ci, v))


@needs_ref
@pytest.mark.parametrize("M,K,avg,long_row", [(300, 9000, 12, None), (1, 5, 3, None), (70, 4096, 5, None),
                                              (1000, 4097, 3, None), (5, 20000, 50, None), (64, 64, 0, None),
                                              (129, 700, 9, 650), (2000, 300, 30, None)])
def test_random_images_decode_to_the_csr(M, K, avg, long_row):
    rp, ci, v = random_csr(M, K, avg, M + K, np.float32, long_row=long_row)
    v[::7] = -0.0  # a value whose bits matter: -0.0 survives, like in the loader
    ptr, imgs, _ = oracle.ref_build_images(M, K, rp, ci, v)
    assert same_csr(sx.images_decode_A(ptr, imgs, M, K), (rp, ci, v))


@needs_ref
def test_decoder_thread_count_does_not_matter(monkeypatch):
    rp, ci, v = random_csr(3000, 10000, 40, 5, np.float32)
    ptr, imgs, _ = oracle.ref_build_images(3000, 10000, rp, ci, v)
    outs = []
    for t in ("1", "3", "64"):
        monkeypatch.setenv("SX_HOST_THREADS", t)
        outs.append(sx.images_decode_A(ptr, imgs, 3000, 10000))
    assert same_csr(outs[0], (rp, ci, v)) and same_csr(outs[1], outs[0]) and same_csr(outs[2], outs[0])


def test_decoder_rejects_what_the_packer_cannot_have_written():
    g = np.load(os.path.join(GOLDEN, "images_small.npz"))
    M, K, _ = g["p_dims"].tolist()
    ptr = g["p_ptr"].copy()
    imgs = [g[f"p_A{c}"].copy() for c in range(8)]
    with pytest.raises(sx.SextansError, match="row >= M"):
        sx.images_decode_A(ptr, imgs, 64, K)          # rows 64..69 exist in the images
    with pytest.raises(sx.SextansError, match="column windows"):
        sx.images_decode_A(ptr, imgs, M, K + 4096)    # NUM_ITE no longer fits K
    bad = ptr.copy()
    bad[1] = bad[2] + 1
    with pytest.raises(sx.SextansError, match="decreases"):
        sx.images_decode_A(bad, imgs, M, K)
    bad = ptr.copy()
    bad[0] = 1
    with pytest.raises(sx.SextansError, match=r"ptr\[0\]"):
        sx.images_decode_A(bad, imgs, M, K)
    with pytest.raises(ValueError):
        sx.images_decode_A(ptr, [im[:8] for im in imgs], M, K)
    # a column of the last window beyond K
    w = imgs[0].copy()
    j = int(ptr[1]) * 8  # first slot of window 1 (columns 4096..): word of PE 0
    w[j] = (np.uint64(4000) << np.uint64(50)) | np.uint64(0x3F800000)
    with pytest.raises(sx.SextansError, match="column >= K"):
        sx.images_decode_A(ptr, [w] + imgs[1:], M, K)


def test_bubbles_are_skipped():
    # one window, one slot, every word a bubble except PE 9's (channel 1, word bitrev3(1) = 4)
    ptr = np.array([0, 1], dtype=np.int32)
    imgs = [np.full(8, np.uint64(0x3FFFF) << np.uint64(32), dtype=np.uint64) for _ in range(8)]
    val = np.array([2.5], dtype=np.float32).view(np.uint32)[0]
    imgs[1][4] = (np.uint64(17) << np.uint64(50)) | (np.uint64(2) << np.uint64(32)) | np.uint64(val)
    rp, ci, v = sx.images_decode_A(ptr, imgs, 200, 100)
    row = 2 * 64 + 9
    assert ci.tolist() == [17] and v.tolist() == [2.5]
    assert rp[row] == 0 and rp[row + 1] == 1 and rp[-1] == 1


@pytest.mark.parametrize("M,K,N", [(45, 37, 24), (16, 8, 8), (4704, 4704, 16), (1, 1, 8), (33, 130, 64)])
def test_dense_image_layouts_round_trip(M, K, N):
    rng = np.random.default_rng(M * 1000 + K)
    B = rng.uniform(-1, 1, K * N).astype(np.float32)
    Cm = rng.uniform(-1, 1, M * N).astype(np.float32)
    L = sx.lib()
    bi, ci = oracle.pack_B_images(B, K, N), oracle.pack_C_images(Cm, M, N)
    assert L.sx_images_B_floats(K, N) <= bi[0].size and L.sx_images_C_floats(M, N) <= ci[0].size
    assert L.sx_images_B_floats(K, N) == (K + 7) // 8 * 8 * 2 * (N // 8)   # host.cpp:154-157
    assert L.sx_images_C_floats(M, N) == (M + 15) // 16 * 16 * (N // 8)    # host.cpp:176
    assert np.array_equal(sx.images_decode_B(bi, K, N), B)
    assert np.array_equal(sx.images_decode_C(ci, M, N), Cm)
    # encode: rows 0..M-1 as the host's read-back finds them, the pad rows of the last
    # 16-row word as alpha*0 + beta*C_in, nothing beyond
    cin = oracle.pack_C_images(Cm, M, N, fill_pad=0.5)
    out = [np.full_like(x, 7.0) for x in cin]
    a, b = np.float32(0.85), np.float32(-2.06)
    sx.images_encode_C(Cm, M, N, a, b, cin, out)
    assert np.array_equal(oracle.unpack_C_images(out, M, N), Cm)
    used = L.sx_images_C_floats(M, N)
    exp = oracle.pack_C_images(Cm, M, N, fill_pad=a * np.float32(0) + b * np.float32(0.5))
    for o, e in zip(out, exp):
        assert np.array_equal(o[:used], e[:used]) and (o[used:] == 7.0).all()


def test_dataflow_model_on_golden_images_equals_the_csr_oracle(golden):
    """The accelerator's dataflow, modelled on its own images in hardware order
    (oracle.sextans_images: src/sextans.cpp:285-570), gives cpu_spmm_CSR's result bit for
    bit -- the property that lets cpu_spmm_CSR stand in for the kernel."""
    g = np.load(os.path.join(GOLDEN, "images_small.npz"))
    for tag in golden["image_cases"]:
        M, K, _ = g[tag + "_dims"].tolist()
        for N, alpha, beta in ((8, 0.85, -2.06), (24, -1.25, 0.5)):
            B, Cin = random_dense(M, K, N, 17, np.float32)
            P_N, au, bu = oracle.pack_scalars(N, 2, alpha, beta)
            cin = oracle.pack_C_images(Cin, M, N, fill_pad=0.25)
            co = oracle.sextans_images(g[tag + "_ptr"], [g[f"{tag}_A{c}"] for c in range(8)],
                                       oracle.pack_B_images(B, K, N), cin, M, K, P_N, au, bu)
            ref = oracle.spmm_csr(M, N, K, g[tag + "_rowptr"], g[tag + "_colidx"], g[tag + "_val"],
                                  np.float32(alpha), B, np.float32(beta), Cin.copy())
            assert np.array_equal(oracle.unpack_C_images(co, M, N).view(np.uint32), ref.view(np.uint32)), tag
            # pad rows of the last 16-row word: alpha*0 + beta*pad, nothing beyond
            pad = np.float32(alpha) * np.float32(0) + np.float32(beta) * np.float32(0.25)
            exp = oracle.pack_C_images(ref, M, N, fill_pad=pad)
            used = sx.lib().sx_images_C_floats(M, N)
            for o, e in zip(co, exp):
                assert np.array_equal(o[:used].view(np.uint32), e[:used].view(np.uint32)) and not o[used:].any()


@needs_ref
@pytest.mark.parametrize("name,N", [("nasa4704", 16), ("nasa4704", 8), ("pcrystk02", 8)])
def test_dataflow_model_equals_the_compiled_reference_on_suitesparse(name, N):
    M, K, nnz, rp, ci, v = oracle.ref_load_csr(mtx_path(name))
    v = (1.0 + 0.001 * (np.arange(nnz) % 97)).astype(np.float32)
    ptr, imgs, _ = oracle.ref_build_images(M, K, rp, ci, v)
    B, Cin = random_dense(M, K, N, 3, np.float32)
    P_N, au, bu = oracle.pack_scalars(N, 1, 0.85, -2.06)
    co = oracle.sextans_images(ptr, imgs, oracle.pack_B_images(B, K, N), oracle.pack_C_images(Cin, M, N),
                               M, K, P_N, au, bu)
    ref = Cin.copy()
    oracle.ref_spmm_csr(M, N, K, rp, ci, v, 0.85, B, -2.06, ref)   # the reference's own cpu_spmm_CSR
    assert np.array_equal(oracle.unpack_C_images(co, M, N).view(np.uint32), ref.view(np.uint32))


def test_scalar_packing_matches_the_host():
    P_N, au, bu = oracle.pack_scalars(16, 3, 0.85, -2.06)
    assert P_N == (3 << 16) | 16
    assert np.array([au, bu], dtype=np.int32).view(np.float32).tolist() == [np.float32(0.85), np.float32(-2.06)]


def test_unmodified_reference_host_builds_against_the_engine_and_fails_loudly_without_a_gpu():
    """oracle/_ref/sextans_ref_host is src/sextans-host.cpp, unmodified, compiled with
    include/tapa_compat + integration/sextans_kernel_b200.cpp.  Without a device it must
    get as far as the device call and then stop with the engine's error -- no CPU fallback."""
    import torch
    exe = os.path.join(ROOT, "oracle", "_ref", "sextans_ref_host")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/sextans_ref_host not built")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (covered by the gpu test)")
    r = subprocess.run([exe, mtx_path("nasa4704"), "16"], capture_output=True, text=True, timeout=120)
    assert "launch kernel" in r.stdout and "Success" not in r.stdout
    assert r.returncode == 1 and "SX_ERR_NO_DEVICE" in r.stderr
