"""CPU tier: the host side of mm_gemm_fwd (tile width from the cost model, CTA pairs / cta_group::2, rasterisation group,
stream-K tail, argument checks) through mm_gemm_plan — the same dispatcher run without a launch.  No GPU involved: the
library assumes 148 SMs when no device is visible.  The decisions asserted here are the ones DESIGN.md §3 / §6 document."""
import ctypes as C
import random

import pytest

from macaw_llm_b200 import _lib, ops

E, I, V = 4096, 11008, 32000
SMS = 148


def plan(**kw):
    return ops.gemm_plan(**kw)


def test_llama_gemms_at_batch_32_run_as_cta_group2_pairs():
    M = 32 * 528
    for N, K, epi in ((3 * E, E, ops.EPI_ROPE), (E, E, ops.EPI_STD), (2 * I, E, ops.EPI_SWIGLU), (E, I, ops.EPI_STD),
                      (V, E, ops.EPI_STD)):
        p = plan(M=M, N=N, K=K, epi=epi, fp16=True, streamk=True)
        assert (p["block_n"], p["pairs"], p["grid"], p["workers"]) == (256, 2, SMS, SMS // 2), (N, K, p)
        assert p["units"] == (M // 256) * ((N + 255) // 256) and p["streamk_tiles"] == 0  # stream-K never rides pairs
        assert p["smem_bytes"] <= 227 * 1024 and p["vectorised_epilogue"] == 1
    # rasterisation group ~ 32 MiB of A rows: 16 pairs at K = 4096, 6 at K = 11008
    assert plan(M=M, N=E, K=E)["group_m"] == 16 and plan(M=M, N=E, K=I)["group_m"] == 6


def test_short_k_gemms_stay_single_cta():
    # CLIP / Whisper layers (K = 1024 / 512): 8 - 16 k-blocks do not amortise a pair's fill (measured slower)
    for M, N, K in ((8224, 3072, 1024), (8224, 4096, 1024), (48000, 1536, 512), (48000, 2048, 512)):
        p = plan(M=M, N=N, K=K, fp16=True)
        assert p["pairs"] == 0 and p["workers"] == SMS and p["grid"] == SMS, p
    assert plan(M=8224, N=1024, K=4096, fp16=True)["pairs"] == 2  # fc2: K = 4096 -> pairs


def test_per_gpu_batch_of_the_8_gpu_run_odd_m_tiles():
    """M = 2112 = 17 M tiles: pairs only where the pair schedule needs no more waves than the single-CTA one."""
    M = 4 * 528
    qkv = plan(M=M, N=3 * E, K=E, epi=ops.EPI_ROPE, fp16=True, streamk=True)
    assert (qkv["pairs"], qkv["units"], qkv["waves"]) == (2, 9 * 48, 6)         # 816 single tiles would also be 6 waves
    gu = plan(M=M, N=2 * I, K=E, epi=ops.EPI_SWIGLU, fp16=True, streamk=True)
    assert (gu["pairs"], gu["units"], gu["waves"]) == (0, 17 * 86, 10)          # pairs would need 11 waves
    head = plan(M=M, N=V, K=E, fp16=True, streamk=True)
    assert head["pairs"] == 0 and head["streamk_tiles"] == (17 * 125) % SMS == 53 and head["grid"] == SMS
    assert plan(M=M, N=V, K=E, fp16=True, streamk=False)["streamk_tiles"] == 0  # opt-in per launch through the workspace


def test_policy_switches():
    lib = _lib.load()
    M = 32 * 528
    prev = lib.mm_gemm_cg2_mode(0)
    try:
        assert plan(M=M, N=E, K=E)["pairs"] == 1  # round 1's scheme: multicast pairs of cta_group::1 MMAs
    finally:
        lib.mm_gemm_cg2_mode(prev)
    assert plan(M=M, N=E, K=E)["pairs"] == 2
    prev = lib.mm_gemm_streamk_mode(0)
    try:
        assert plan(M=4 * 528, N=V, K=E, streamk=True)["streamk_tiles"] == 0
    finally:
        lib.mm_gemm_streamk_mode(prev)
    prev = lib.mm_gemm_streamk_mode(2)  # whenever the schedule allows
    try:
        p = plan(M=4 * 528, N=E, K=E, streamk=True)
        assert p["pairs"] == 2 and p["streamk_tiles"] == 0  # ... which excludes pair launches
        assert plan(M=1028, N=4096, K=1024, streamk=True)["streamk_tiles"] == 0   # needs more than one full wave
        assert plan(M=8224, N=3072, K=1024, streamk=True)["streamk_tiles"] == 780 % SMS
    finally:
        lib.mm_gemm_streamk_mode(prev)


def test_thin_decode_gemms_use_narrow_tiles():
    # swapped operands (c_trans): the weight rows fill the 128-row MMA tile, the 8 token rows are the N extent
    p = plan(M=E, N=8, K=E, c_trans=True)
    assert p["block_n"] == 32 and p["pairs"] == 0 and p["units"] == E // 128 and p["grid"] == E // 128
    p = plan(M=E, N=8, K=E // 4, batch=4, c_fp32=True)  # split-K of 4 through the batch dimension: 128 units
    assert p["units"] == 4 * (E // 128) and p["grid"] == 128


def test_schedule_invariants_random_shapes():
    rng = random.Random(7)
    for _ in range(300):
        M = rng.choice([1, 7, 128, 129, 257, 1028, 2112, 4224, 6000, 16896]) + rng.choice([0, 0, 8, 64])
        N = rng.choice([8, 64, 96, 512, 768, 1024, 1536, 3072, 4096, 12288, 22016, 32000])
        K = rng.choice([64, 256, 512, 768, 1024, 2048, 4096, 11008])
        batch = rng.choice([1, 1, 1, 2, 16])
        b_mn = rng.random() < 0.2
        a_mn = b_mn and rng.random() < 0.5
        if a_mn:
            M = (M + 7) // 8 * 8  # lda = M must be a multiple of 8
        if b_mn:
            N = (N + 7) // 8 * 8
        p = plan(M=M, N=N, K=K, batch=batch, b_mn_major=b_mn, a_mn_major=a_mn, streamk=rng.random() < 0.5)
        assert p["block_n"] in (32, 64, 128, 256) and (not b_mn or p["block_n"] >= 64)
        assert p["m_tiles"] == (M + 127) // 128 and p["n_tiles"] == (N + p["block_n"] - 1) // p["block_n"]
        assert p["k_blocks"] == (K + 63) // 64
        m_units = (p["m_tiles"] + 1) // 2 if p["pairs"] else p["m_tiles"]
        assert p["units"] == batch * m_units * p["n_tiles"]
        assert p["workers"] == (SMS // 2 if p["pairs"] else SMS)
        assert p["waves"] == -(-p["units"] // p["workers"]) and 0 < p["grid"] <= SMS
        assert p["grid"] % 2 == 0 or not p["pairs"]
        assert not (p["pairs"] and (p["block_n"] != 256 or p["k_blocks"] < 32 or a_mn))
        assert 0 <= p["streamk_tiles"] < SMS and not (p["streamk_tiles"] and p["pairs"])
        assert p["smem_bytes"] <= 227 * 1024 and p["group_m"] >= 2


def test_argument_checks_raise_with_a_message():
    lib = _lib.load()
    fake = 1 << 20

    def rc(**over):
        kw = dict(M=256, N=256, K=256, batch=1, batch2=1, A=fake, lda=256, B=fake, ldb=256, C=fake, ldc=256, alpha=1.0)
        kw.update(over)
        a = _lib.GemmArgs(**kw)
        out = _lib.GemmPlan()
        r = lib.mm_gemm_plan(C.byref(a), C.byref(out))
        return r, _lib.last_error()

    assert rc()[0] == 0
    r, msg = rc(a_fp16=1, b_fp16=0)
    assert r != 0 and "mixed f16 x bf16" in msg  # sm_100a faults on a mixed-format tcgen05.mma (measured)
    r, msg = rc(lda=250)
    assert r != 0 and "multiples of 8" in msg
    r, msg = rc(A=fake + 2)
    assert r != 0 and "16-byte aligned" in msg
    r, msg = rc(epi=ops.EPI_ROPE)
    assert r != 0 and "RoPE" in msg          # no cos / sin tables
    r, msg = rc(a_mn_major=1)
    assert r != 0 and "MN-major A" in msg    # needs MN-major B as well
    r, msg = rc(M=0)
    assert r != 0 and "bad shape" in msg
    with pytest.raises(RuntimeError):
        ops.gemm_plan(M=256, N=100, K=256, epi=ops.EPI_SWIGLU)  # SwiGLU epilogue needs N % 64 == 0
