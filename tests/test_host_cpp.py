"""The compiled (C++) host layer -- csrc/host/ntn_host.hpp through its C-ABI test hooks -- against the oracle,
bit for bit: java.util.Random stream, entity-name parsing -> forward/backward word maps (with the reference's
malformed-name quirks), batch-job order and corrupt indices.  No GPU needed."""
import ctypes as C
C_ = C
import os

import numpy as np
import pytest

from neural_tensor_network_for_kb_completion_b200 import _capi, synth
from oracle import data_ref
from oracle.java_random import JavaRandom


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _capi.load()


def jr(lib, seed, bound, n, m=0):
    ints, dbl = np.zeros(max(n, 1), np.int32), np.zeros(max(m, 1), np.float64)
    assert lib.ntn_host_java_random(seed, bound, n, ints.ctypes.data, m, dbl.ctypes.data) == 0
    return ints[:n], dbl[:m]


def test_cpp_java_random_known_answers_and_stream(lib):
    assert jr(lib, 42, 0, 1)[0][0] == -1170105035 and jr(lib, 0, 0, 1)[0][0] == -1155484576
    assert jr(lib, 42, 100, 5)[0].tolist() == [30, 63, 48, 84, 70]
    assert jr(lib, 42, 75043, 8)[0].tolist() == [35870, 25511, 45555, 64931, 40108, 3288, 5558, 26082]
    assert jr(lib, 42, 0, 0, 3)[1].tolist() == [0.7275636800328681, 0.6832234717598454, 0.30871945533265976]
    for seed, bound in ((7, 38696), (123456789, 1024), (-5, 3), (99, (1 << 30) + 1)):
        ints, dbl = jr(lib, seed, bound, 3000, 100)
        r = JavaRandom(seed)
        assert ints.tolist() == [r.next_int(bound) for _ in range(3000)]
        assert dbl.tolist() == [r.next_double() for _ in range(100)]


def test_cpp_entity_word_maps_match_oracle_including_quirks(lib):
    kb = synth.make_kb(800, 5, 10, 120, 4, oov_frac=0.15, seed=5)
    names = list(kb.entity_names)
    names[3], names[9], names[17], names[40] = "__w1__1", "__2", "__w5_w6_w7_w8_3", "__a_"
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    want = data_ref.build_entity_word_maps(names, vocab)
    Ne, cap = len(names), 8 * len(names)
    fo, bo, lb = np.zeros(Ne + 1, np.int32), np.zeros(Ne + 1, np.int32), np.zeros(Ne, np.int32)
    fi, bi = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    nf, nb = C.c_int64(), C.c_int64()
    rc = lib.ntn_host_entity_word_maps("\n".join(names).encode(), "\n".join(kb.vocab_words).encode(), fo.ctypes.data,
                                       fi.ctypes.data, bo.ctypes.data, bi.ctypes.data, lb.ctypes.data, cap, C.byref(nf), C.byref(nb))
    assert rc == 0
    got = (fo, fi[:nf.value], bo, bi[:nb.value], lb)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert not np.array_equal(fo, bo)                       # the malformed names make the two lists differ


def test_cpp_batch_job_matches_oracle_across_consecutive_jobs(lib):
    kb = synth.make_kb(500, 4, 300, 50, 4, seed=6)
    train = np.ascontiguousarray(kb.train, np.int32)
    B, Cc, n = 120, 7, 120 * 7
    out = [np.zeros(n, np.int32) for _ in range(4)]
    rc = lib.ntn_host_generate_batch(train.ctypes.data, len(train), B, Cc, kb.Ne, 42, 3, *(o.ctypes.data for o in out))
    assert rc == 0
    r = JavaRandom(42)
    for _ in range(3):                                      # the Random stream continues from job to job
        want = data_ref.generate_batch_job(train[:, 0], train[:, 1], train[:, 2], B, Cc, kb.Ne, r)
    for a, b in zip(out, want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 17, 4096, 4099, 70001, 300007])
@pytest.mark.parametrize("shift", [0, 1, 3])
def test_staging_conversions_on_ragged_unaligned_ranges(n, shift):
    """double[] theta -> float staging and float gradient -> double[] (Util.java:49-55 rounding): the threaded,
    non-temporal-store loops must equal the plain casts for every length and destination alignment."""
    lib = _capi.load()
    rng = np.random.default_rng(n + shift)
    src64 = np.empty(n + shift + 8, np.float64)[shift:shift + n]
    src64[:] = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, size=n)
    dst32 = np.full(n + shift + 8, np.float32(7.0))
    assert lib.ntn_host_convert_f64_f32(src64.ctypes.data, dst32[shift:].ctypes.data, n) == 0
    with np.errstate(over="ignore"):
        np.testing.assert_array_equal(dst32[shift:shift + n], src64.astype(np.float32))
    assert np.all(dst32[:shift] == 7.0) and np.all(dst32[shift + n:] == 7.0)
    src32 = dst32[shift:shift + n].copy()
    dst64 = np.full(n + shift + 8, 7.0)
    assert lib.ntn_host_convert_f32_f64(src32.ctypes.data, dst64[shift:].ctypes.data, n) == 0
    np.testing.assert_array_equal(dst64[shift:shift + n], src32.astype(np.float64))
    assert np.all(dst64[:shift] == 7.0) and np.all(dst64[shift + n:] == 7.0)


@pytest.mark.parametrize("n,seed", [(0, 1), (1, 1), (2, 5), (17, 42), (1000, 7)])
def test_collections_shuffle_three_ways(n, seed):
    """The shuffle commented out at DataFactory.java:118, re-enabled with a seeded stream: oracle restatement of
    java.util.Collections.shuffle == C++ host == Python DataFactory mirror."""
    from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
    ref = data_ref.collections_shuffle(n, JavaRandom(seed))
    assert sorted(ref.tolist()) == list(range(n))
    lib = _capi.load()
    perm = np.empty(max(n, 1), np.int32)
    assert lib.ntn_host_collections_shuffle(seed, n, perm.ctypes.data) == 0
    np.testing.assert_array_equal(perm[:n], ref)
    f = DataFactory(4, 2, 3, seed=0, shuffle_seed=seed)
    f.trainingTripples = np.stack([np.arange(n), np.zeros(n, int), np.arange(n)], 1).astype(np.int32).reshape(-1, 3)
    got = f.shuffleTrainingTripples()
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(f.trainingTripples[:, 0], ref)


def test_device_sampler_restatement_is_uniform_and_counter_based():
    """oracle/data_ref.py: device_corrupt_sample is a pure function of (seed, job, column, attempt)."""
    j = np.arange(200000)
    a = data_ref.device_corrupt_sample(42, 3, j, 1000)
    assert a.min() >= 0 and a.max() < 1000
    cnt = np.bincount(a, minlength=1000)
    assert cnt.min() > 120 and cnt.max() < 290                      # 200 +- 5.7 sigma
    np.testing.assert_array_equal(a[1234:1240], data_ref.device_corrupt_sample(42, 3, j[1234:1240], 1000))
    assert not np.array_equal(a, data_ref.device_corrupt_sample(42, 4, j, 1000))
    assert not np.array_equal(a, data_ref.device_corrupt_sample(43, 3, j, 1000))
    assert not np.array_equal(a, data_ref.device_corrupt_sample(42, 3, j, 1000, attempt=1))


@pytest.mark.parametrize("d", [1, 2, 3, 7, 13, 25, 26, 33, 64, 65, 100, 257, 4096, 65537, 2**31 - 1, 2**31, 2**32 - 1])
def test_invariant_division_of_the_flat_thread_index(d):
    """The gather kernels divide their flat 32-bit thread index by the run-time row width with precomputed magic
    numbers (ntn_kernels.cu: FastDiv); the quotient must be exact for every 32-bit dividend."""
    lib = _capi.load()
    rng = np.random.default_rng(d % 1000)
    edge = np.array([0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2**31 - 1, 2**31, 2**32 - 1], dtype=np.int64) % (2**32)
    n = np.concatenate([edge, rng.integers(0, 2**32, size=20000, dtype=np.int64),
                        (rng.integers(0, 2**32 // d + 1, size=2000, dtype=np.int64) * d) % (2**32)]).astype(np.uint32)
    q = np.empty_like(n)
    assert lib.ntn_host_fastdiv(d, len(n), n.ctypes.data, q.ctypes.data) == 0
    np.testing.assert_array_equal(q, (n.astype(np.uint64) // np.uint64(d)).astype(np.uint32))


def test_theta_checkpoints_round_trip_bit_exact(tmp_path):
    """Run_NTN.java:133-139 writes theta as comma-separated text; the resume (commented out at :108) reads it back.
    Both hosts, both formats (text and the binary fast path), and across hosts: every FP32 value comes back bit for bit."""
    import ctypes as C
    from neural_tensor_network_for_kb_completion_b200 import _capi, run_ntn
    lib = _capi.load()
    rng = np.random.default_rng(0)
    th = np.concatenate([rng.standard_normal(5000) * 10.0 ** rng.integers(-30, 30, 5000),
                         [0.0, -0.0, 1e-45, 3.4028235e38, -1.17549435e-38, 0.1, 1.0 / 3.0]]).astype(np.float32)
    th64 = th.astype(np.float64)
    bits = th.view(np.uint32)
    for binary in (0, 1):
        path = str(tmp_path / ("c_theta.bin" if binary else "c_theta.txt"))
        back = np.empty_like(th64)
        rc = lib.ntn_host_theta_checkpoint_roundtrip(path.encode(), binary, th64.ctypes.data_as(C.c_void_p), th64.size,
                                                     back.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert np.array_equal(back.astype(np.float32).view(np.uint32), bits), binary
        # the Python host reads what the C++ host wrote
        assert np.array_equal(run_ntn.read_theta(path).astype(np.float32).view(np.uint32), bits)
    # ... and the other way round: Python writes, C++ reads (through its round trip helper's reader) and Python reads
    ptxt, pbin = str(tmp_path / "py_theta.txt"), str(tmp_path / "py_theta.bin")
    run_ntn.save_checkpoint(th64, ptxt)
    assert os.path.exists(pbin)
    for path in (ptxt, pbin):
        assert np.array_equal(run_ntn.read_theta(path).astype(np.float32).view(np.uint32), bits)
    assert open(ptxt).read().count(",") == th.size - 1 and "\n" not in open(ptxt).read()      # one comma-separated line


def test_native_corrupt_side_drawer_continues_the_seeded_stream(lib):
    """ntn_host_sides_draw (the ntn_sides_fn a non-compiled host hands to ntn_lbfgs_minimize) against the oracle's
    java.util.Random: one nextDouble() > 0.5 per PRESENT relation in increasing r (NTN.java:120,146), state carried."""
    from oracle.java_random import JavaRandom as RefRandom
    R = 13
    present = np.array([1, 1, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, 1], np.uint8)
    ref = RefRandom(7)
    from neural_tensor_network_for_kb_completion_b200.data_factory import JavaRandom as PyRandom
    pr = PyRandom(7)
    st = np.zeros(2 + R, np.uint64)
    st[0], st[1] = pr.s, R
    st[2:] = present
    out = np.zeros(R, np.uint8)
    for _ in range(50):
        lib.ntn_host_sides_draw(st.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        want = np.array([(1 if ref.next_double() > 0.5 else 0) if present[r] else 0 for r in range(R)], np.uint8)
        assert np.array_equal(out, want)
    # the Python mirror adopts the advanced state and stays on the same stream
    pr.s = int(st[0])
    assert pr.nextDouble() == ref.next_double()


@pytest.mark.parametrize("Nw,dq,C", [(67447, 100, 1), (67447, 100, 2), (67447, 100, 4), (420, 20, 3), (5, 8, 4), (500000, 256, 8)])
def test_word_pull_chunk_layout_covers_every_word_once(lib, Nw, dq, C):
    """The data-parallel exchange cuts the word pull into chunks (csrc/ntn_kernels.cu: ntn_word_chunk): contiguous, whole
    words, every word in exactly one chunk, block-partial slots back to back (the regulariser sum reads them all)."""
    out = np.zeros(4 * C, np.int32)
    total = lib.ntn_host_word_chunks(Nw, dq, C, out.ctypes.data_as(C_.c_void_p))
    ch = out.reshape(C, 4)
    assert ch[0, 0] == 0 and ch[:, 1].sum() == Nw
    assert np.array_equal(ch[1:, 0], np.cumsum(ch[:-1, 1]))
    assert np.array_equal(ch[:, 2], np.r_[0, np.cumsum(ch[:-1, 3])])
    assert total == ch[:, 3].sum()
    cpr = dq // 4
    assert np.array_equal(ch[:, 3], -(-((ch[:, 1] + 1) // 2 * cpr) // 256))      # one thread per (word pair, float4 chunk)
    assert ch[:, 1].max() - ch[:, 1].min() <= 1 or C > Nw
