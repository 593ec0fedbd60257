"""DataFactory restated (oracle; test infra) -- entity-name parsing, entity->word
maps, the batch / corrupt sampler and the entity-vector build.

Follows /root/reference/src/DataFactory.java; every function cites its lines.
Pure Python / NumPy on purpose: it is the independent derivation the C++ host
(`csrc/host/`) and the CUDA kernels are checked against.
"""
from __future__ import annotations

import numpy as np

from .java_random import JavaRandom


# --------------------------------------------------------------------------
# java.lang.String helpers with Java semantics
# --------------------------------------------------------------------------
def java_split(s: str, sep: str = "_") -> list[str]:
    """``s.split(sep)`` for a single-literal-character separator.

    Java semantics (String.split(regex) == split(regex, 0)): a positive-width
    match at index 0 yields a leading empty string; trailing empty strings are
    removed; if nothing matches the result is [s].  Python's str.split keeps
    the trailing empties, so strip them here.
    """
    parts = s.split(sep)
    while len(parts) > 1 and parts[-1] == "":
        parts.pop()
    if len(parts) == 1 and parts[0] == "" and s != "":
        # s consisted only of separators: Java returns an empty array
        return []
    return parts


def entity_name_clear(raw: str) -> str:
    """DataFactory.java:397-402 (also :162-167, :655-660):
    ``raw.substring(2, raw.lastIndexOf("_"))`` with the catch-all fallback
    ``raw.substring(2)`` when the first form throws."""
    last = raw.rfind("_")
    # Java substring(begin, end) throws if begin > end or end < 0 or end > len
    if last >= 2 and len(raw) >= 2:
        return raw[2:last]
    if len(raw) < 2:
        # substring(2) would throw again -> propagate like the reference would
        raise ValueError(f"entity name too short for the reference parser: {raw!r}")
    return raw[2:]


def entity_forward_words(raw: str) -> list[str]:
    """Words averaged in the forward pass, DataFactory.java:405-434:
    multi-word branch iff the clear name contains '_' (then split("_")),
    else the clear name itself."""
    name = entity_name_clear(raw)
    if "_" in name:
        return java_split(name, "_")
    return [name]


def entity_length(raw: str) -> int:
    """DataFactory.java:583-595: ``raw.split("_").length - 3``."""
    return len(java_split(raw, "_")) - 3


def word_index(word: str, vocab: dict[str, int]) -> int:
    """vocabWordNum.get(word) with the exception-driven "unknown" fallback
    (DataFactory.java:411-416, :428-433, :666-680)."""
    i = vocab.get(word)
    if i is None:
        i = vocab["unknown"]
    return i


def get_word_indexes(raw: str, vocab: dict[str, int]) -> list[int]:
    """DataFactory.java:646-684.  The array has ``entityLength`` slots (1 if that
    is 0) and is filled from the *forward* word list; unfilled slots keep Java's
    default 0, so for malformed names the backward list can differ from the
    forward one."""
    n = entity_length(raw)
    out = [0] * (n if n != 0 else 1)
    words = entity_forward_words(raw)
    if len(words) > len(out):
        raise IndexError(f"reference would throw ArrayIndexOutOfBounds for {raw!r}")
    for j, w in enumerate(words):
        out[j] = word_index(w, vocab)
    return out


def build_entity_word_maps(names: list[str], vocab: dict[str, int]):
    """Forward CSR (words averaged, with multiplicity), backward CSR (words the
    gradient is added to) and the backward divisor per entity.

    Returns (fwd_off, fwd_idx, bwd_off, bwd_idx, len_bwd) as int32 arrays."""
    fwd_off = [0]
    fwd_idx: list[int] = []
    bwd_off = [0]
    bwd_idx: list[int] = []
    len_bwd = []
    for raw in names:
        fw = [word_index(w, vocab) for w in entity_forward_words(raw)]
        fwd_idx.extend(fw)
        fwd_off.append(len(fwd_idx))
        bw = get_word_indexes(raw, vocab)
        bwd_idx.extend(bw)
        bwd_off.append(len(bwd_idx))
        len_bwd.append(entity_length(raw))
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    return i32(fwd_off), i32(fwd_idx), i32(bwd_off), i32(bwd_idx), i32(len_bwd)


# --------------------------------------------------------------------------
# batch job
# --------------------------------------------------------------------------
def generate_batch_job(train_e1, train_rel, train_e2, batch_size: int, corrupt_size: int,
                       num_entities: int, rng: JavaRandom):
    """DataFactory.java:116-136: h-major, always the first ``batch_size`` training
    triples (shuffle commented out at :118), e3 = rand.nextInt(numOfentities)."""
    n = batch_size * corrupt_size
    e1 = np.empty(n, np.int32)
    rel = np.empty(n, np.int32)
    e2 = np.empty(n, np.int32)
    e3 = np.empty(n, np.int32)
    j = 0
    for _h in range(corrupt_size):
        for i in range(batch_size):
            e3[j] = rng.next_int(num_entities)
            e1[j] = train_e1[i]
            rel[j] = train_rel[i]
            e2[j] = train_e2[i]
            j += 1
    return e1, rel, e2, e3


def collections_shuffle(n: int, rng: JavaRandom) -> np.ndarray:
    """java.util.Collections.shuffle(list, rnd) (JDK spec: ``for (int i = size; i > 1; i--) swap(list, i-1,
    rnd.nextInt(i))``) as the permutation it applies: result[p] = original index now at position p.  This is the
    shuffle the reference has commented out at DataFactory.java:118 ("TODO wieder zulassen"), with a seeded stream."""
    perm = np.arange(n, dtype=np.int64)
    for i in range(n, 1, -1):
        j = rng.next_int(i)
        perm[i - 1], perm[j] = perm[j], perm[i - 1]
    return perm


_M64 = (1 << 64) - 1


def device_corrupt_mix(seed: int, job: int, j, attempt: int = 0) -> np.ndarray:
    """The counter-based 64-bit mix of the device sampler (csrc/batch_index.cu: bs_mix), vectorised over columns j."""
    with np.errstate(over="ignore"):
        j = np.asarray(j, dtype=np.uint64)
        z = np.uint64((seed + 0x9E3779B97F4A7C15 * (1 + job)) & _M64)
        z = z ^ (j * np.uint64(0xBF58476D1CE4E5B9))
        z = z + np.uint64((attempt * 0x94D049BB133111EB) & _M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def device_corrupt_sample(seed: int, job: int, j, num_entities: int, attempt: int = 0) -> np.ndarray:
    """e3 = ((mix >> 32) * Ne) >> 32 (multiply-shift range reduction of the high 32 bits)."""
    hi = device_corrupt_mix(seed, job, j, attempt) >> np.uint64(32)
    return ((hi * np.uint64(num_entities)) >> np.uint64(32)).astype(np.int32)


def generate_batch_job_device(train, first: int, batch_size: int, corrupt_size: int, num_entities: int,
                              seed: int, job: int, filter_true: bool = False, max_attempts: int = 16):
    """The device batch-job generator (ntn_generate_batch_dev) restated: h-major columns like
    DataFactory.java:123-129 over training triples [first, first+batch), corrupt entities from the counter-based
    sampler; in filtered mode a draw is repeated (attempt 1..15) while (e1, rel, e3) or (e3, rel, e2) is a training
    triple, the last draw standing if all are rejected."""
    train = np.asarray(train)
    n = batch_size * corrupt_size
    j = np.arange(n, dtype=np.int64)
    t = first + (j % max(batch_size, 1))
    e1, rel, e2 = (train[t, c].astype(np.int32) for c in range(3))
    e3 = device_corrupt_sample(seed, job, j, num_entities, 0)
    if filter_true:
        true = set(map(tuple, train[:, :3].tolist()))
        for col in range(n):
            for att in range(max_attempts):
                c = int(device_corrupt_sample(seed, job, col, num_entities, att))
                e3[col] = c
                if (int(e1[col]), int(rel[col]), c) not in true and (c, int(rel[col]), int(e2[col])) not in true:
                    break
    return e1, rel, e2, e3


def draw_corrupt_sides(rel, num_relations: int, rng: JavaRandom) -> np.ndarray:
    """NTN.java:120,146: one ``Math.random() > 0.5`` per evaluation per *non-empty*
    relation, in increasing r; True => negatives are (e1, e3) ("corrupt tail").
    Stand-in stream: JavaRandom.nextDouble()."""
    present = np.zeros(num_relations, bool)
    present[np.unique(rel)] = True
    out = np.zeros(num_relations, np.uint8)
    for r in range(num_relations):
        if present[r]:
            out[r] = 1 if rng.next_double() > 0.5 else 0
    return out


# --------------------------------------------------------------------------
# entity vectors
# --------------------------------------------------------------------------
def create_entity_vectors(wv: np.ndarray, fwd_off, fwd_idx, dtype=np.float64) -> np.ndarray:
    """DataFactory.java:393-437.  ``wv`` is d x Nw (one column per word).
    Multi-word: zeros + add(...) in word order then div(count) (:407-421);
    single word: the vector itself (:429).  Returns d x Ne."""
    wv = wv.astype(dtype, copy=False)
    d = wv.shape[0]
    ne = len(fwd_off) - 1
    out = np.zeros((d, ne), dtype)
    cnt = np.diff(fwd_off)
    maxlen = int(cnt.max()) if ne else 0
    # accumulate in word order (j = 0, 1, ...) exactly like the reference loop
    for j in range(maxlen):
        sel = np.nonzero(cnt > j)[0]
        out[:, sel] += wv[:, fwd_idx[fwd_off[sel] + j]]
    multi = cnt > 1
    out[:, multi] /= cnt[multi].astype(dtype)
    return out
