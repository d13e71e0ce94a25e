"""Plain-Python models of the parallel formulations the kernels use, checked against the (pinned) serial
restatements in oracle/.  They document WHY the kernels may reorder the reference's serial loops:

  * scan assembly (csrc/assemble.cu): a scan [s, e) between two scan-start nodes is published iff no reset
    position r has s < r <= e; lengths clamp at the holder capacity with the scan's last node in the last slot;
  * standard-node decoder (csrc/decode_formats.cu): the 5-state byte machine as composition of per-byte
    state->state maps (associative, so chunks can be folded independently and scanned);
  * ultra-dense smoothing (csrc/decode_formats.cu): a capsule's effect on `_last_dist_q2` is a table of at most
    nine outcomes selected by its first sample, and a capsule whose nine outcomes agree cuts the chain.

CPU only, small sizes (pure-Python loops)."""
import numpy as np
import pytest


# ---- scan assembly ---------------------------------------------------------------------------------------
def assemble_model(flags, resets, cap):
    """Descriptors (start, stored length, last-node index) of the published scans, the way the kernel finds them."""
    starts = np.flatnonzero(flags & 1)
    resets = np.sort(np.asarray(resets, dtype=np.int64))
    upto = lambda x: int(np.searchsorted(resets, x, side="right"))  # number of reset positions <= x
    out = []
    for s, e in zip(starts[:-1], starts[1:]):
        if upto(e) == upto(s):
            n = int(e - s)
            out.append((int(s), min(n, cap), int(e - 1)))
    return out


@pytest.mark.parametrize("seed", range(12))
def test_scan_assembly_descriptor_rule(oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 1500))
    nodes = np.zeros(n, oracle.NODE_DTYPE)
    nodes["angle_z_q14"] = rng.integers(0, 65536, n)
    nodes["dist_mm_q2"] = np.arange(n) + 1  # identifies the node
    nodes["flag"] = np.where(rng.random(n) < rng.choice([0.01, 0.1, 0.5]), 1, 2)
    resets = np.sort(rng.integers(0, n + 1, int(rng.integers(0, 10)))).astype(np.uint32)
    cap = int(rng.choice([4, 37, 8192]))
    scans, lens, k = oracle.assemble_scans(nodes, resets, cap, 2048)
    model = assemble_model(nodes["flag"], resets, cap)
    assert k == len(model)
    for i, (s, ln, last) in enumerate(model):
        assert lens[i] == ln
        assert (scans[i, : ln - 1]["dist_mm_q2"] == nodes[s: s + ln - 1]["dist_mm_q2"]).all()
        assert scans[i, ln - 1]["dist_mm_q2"] == nodes[last]["dist_mm_q2"]  # capacity rule: last node in the last slot


# ---- standard nodes: byte machine as map composition ---------------------------------------------------------
def byte_map(b):
    t0 = 1 if ((b >> 1) ^ b) & 1 else 0
    t1 = 2 if b & 1 else 0
    return (t0, t1, 3, 4, 0)


def compose(g, f):  # (g o f)(s) = g(f(s))
    return tuple(g[f[s]] for s in range(5))


@pytest.mark.parametrize("seed", range(6))
def test_byte_machine_by_map_composition(oracle, seed):
    rng = np.random.default_rng(seed)
    n = 700
    rec = np.zeros((n, 5), np.uint8)
    rec[:, 0] = (rng.integers(0, 64, n).astype(np.uint8) << 2) | 2
    w = (rng.integers(0, 360 * 64, n).astype(np.uint16) << 1) | 1
    rec[:, 1], rec[:, 2] = w & 0xFF, w >> 8
    rec[:, 3:] = rng.integers(0, 256, (n, 2))
    b = rec.reshape(-1).copy()
    b[rng.choice(len(b), 80, replace=False)] = rng.integers(0, 256, 80)
    b = np.delete(b, rng.choice(len(b), 9, replace=False))
    _, ends, _ = oracle.decode_normal(b)
    # fold chunks of 20 bytes independently, scan the maps, replay each chunk from its entry state
    chunk = 20
    maps = []
    for c0 in range(0, len(b), chunk):
        f = (0, 1, 2, 3, 4)
        for x in b[c0: c0 + chunk]:
            f = compose(byte_map(int(x)), f)
        maps.append(f)
    state, got = 0, []
    for ci, c0 in enumerate(range(0, len(b), chunk)):
        st = state
        for i, x in enumerate(b[c0: c0 + chunk]):
            if st == 4:
                got.append(c0 + i)
            st = byte_map(int(x))[st]
        assert st == maps[ci][state]  # the folded map predicts the chunk's exit state
        state = st
    assert got == ends.tolist()
    # associativity: folding maps pairwise in any grouping gives the same total map
    total = (0, 1, 2, 3, 4)
    for f in maps:
        total = compose(f, total)
    pair = [compose(maps[i + 1], maps[i]) if i + 1 < len(maps) else maps[i] for i in range(0, len(maps), 2)]
    total2 = (0, 1, 2, 3, 4)
    for f in pair:
        total2 = compose(f, total2)
    assert total == total2


# ---- ultra-dense smoothing chain ---------------------------------------------------------------------------------
def ud_samples(cap):
    """(raw distance, scale) of the 64 samples of one 170-byte ultra-dense capsule."""
    out = []
    for pos in range(64):
        cab = cap[10 + 5 * (pos >> 1): 15 + 5 * (pos >> 1)].astype(np.int64)
        lo = int(cab[2] | (cab[3] << 8)) if pos & 1 else int(cab[0] | (cab[1] << 8))
        hi = int(cab[4] >> 4) if pos & 1 else int(cab[4] & 0xF)
        qds = lo | (hi << 16)
        sc = qds & 3
        d = [(qds & 0xFFC) * 2, (qds & 0x1FFC) * 3 + (2046 << 2), (qds & 0x3FFC) * 4 + (8187 << 2),
             (qds & 0x7FFC) * 5 + (24567 << 2)][sc]
        out.append((d, sc))
    return out


def smooth(raw, sc, last):
    return (raw + last) >> 1 if (sc == 0 and last and abs(raw - last) <= 8) else raw


def chain(samples, last):
    for raw, sc in samples:
        last = smooth(raw, sc, last)
    return last


def capsule_table(samples):
    """(first raw, first scale, nine outcomes, constant?) -- what every capsule thread tabulates."""
    r0, sc0 = samples[0]
    outs = [chain(samples[1:], (r0 - 4 + k) if sc0 == 0 else r0) for k in range(9)]
    return r0, sc0, outs, len(set(outs)) == 1


def apply_table(tab, last):
    r0, sc0, outs, _ = tab
    k = 4
    if sc0 == 0 and last and abs(r0 - last) <= 8:
        k = ((r0 + last) >> 1) - (r0 - 4)
    return outs[k]


@pytest.mark.parametrize("near", [False, True])
def test_ultra_dense_chain_tables(oracle, near):
    from test_capsule_oracle_vs_ref import make_capsules

    caps = make_capsules(oracle, 0x86, 60, 50.0, seed=3 if near else 4, near=near)
    tabs = [capsule_table(ud_samples(c)) for c in caps]
    rng = np.random.default_rng(0)
    # 1. a table reproduces the serial chain for any incoming value
    for tab, c in zip(tabs, caps):
        s = ud_samples(c)
        for last in [0, 1, s[0][0], s[0][0] + 8, s[0][0] - 8, s[0][0] + 9, 100000] + rng.integers(0, 9000, 20).tolist():
            if last < 0:
                continue
            assert apply_table(tab, int(last)) == chain(s, int(last))
    # 2. the value entering capsule e: walk back to the nearest constant capsule, apply the tables from there
    serial, last = [], 777
    for c in caps:
        serial.append(last)
        last = chain(ud_samples(c), last)
    for e in range(len(caps)):
        e0 = e - 1
        while e0 >= 0 and not tabs[e0][3]:
            e0 -= 1
        v = tabs[e0][2][0] if e0 >= 0 else 777
        for j in range(e0 + 1, e):
            v = apply_table(tabs[j], v)
        assert v == serial[e]
    assert any(t[3] for t in tabs) or near


# ---- dense / ultra-dense scan-start flags ---------------------------------------------------------------------
K_FULL = 360 << 16


def raw_sync_bits(prev_q8, inc, n):
    """The reference's per-sample test ((cur + inc) % 360deg) < 2 * inc, before the 'not twice in a row' rule."""
    cur, bits = prev_q8 << 8, []
    for _ in range(n):
        bits.append(1 if ((cur + inc) % K_FULL) < (inc << 1) else 0)
        cur += inc
    return bits


def resolve(bits, s_in):
    out, last = [], s_in
    for b in bits:
        s = b & (1 - last)  # sync = (raw ^ last) & raw
        out.append(s)
        last = s
    return out


@pytest.mark.parametrize("n", [40, 64])
def test_no_wrap_shortcut_and_transfer_functions(n):
    """(1) If the first remainder is already >= 2*inc and the capsule cannot wrap, no sample is a scan start
    (the shortcut the decoders take for most capsules).  (2) The flag leaving a capsule is a function of the
    flag entering it; composing those 2-entry functions over capsules equals the serial recurrence."""
    rng = np.random.default_rng(n)
    hits = 0
    for _ in range(4000):
        prev_q8 = int(rng.integers(0, 360 << 8))
        diff_q8 = int(rng.integers(1, 40 << 8)) if rng.random() < 0.8 else int(rng.integers(1, 360 << 8))
        inc = (diff_q8 << 8) // n
        rem = ((prev_q8 << 8) + inc) % K_FULL
        bits = raw_sync_bits(prev_q8, inc, n)
        if rem >= (inc << 1) and rem + (n - 1) * inc < K_FULL:
            hits += 1
            assert not any(bits)
    assert hits > 1000
    # transfer functions
    caps = [raw_sync_bits(int(rng.integers(0, 360 << 8)), int(rng.integers(1, 300000)), n) for _ in range(200)]
    # make some capsules end right after a wrap so that the state matters
    f = [(resolve(b, 0)[-1], resolve(b, 1)[-1]) for b in caps]
    for s0 in (0, 1):
        serial, s = [], s0
        for b in caps:
            serial.append(s)
            s = resolve(b, s)[-1]
        composed, acc = [], (0, 1)  # identity
        for fj in f:
            composed.append(acc[s0])
            acc = (fj[acc[0]], fj[acc[1]])
        assert composed == serial


# ---- the hot path: sort = rank lookup over a presence bitmap ---------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 360, 3200, 32768])
def test_rank_lookup_equals_the_sort_for_tie_free_scans(oracle, n):
    """rank(key) = prefix[key >> 5] + popcount(bits[key >> 5] & below(key)) over the 65536-bit presence map is
    the position std::sort gives the node, and `popcount(map) != measured count` detects duplicate keys."""
    nodes = oracle.synth_batch(90 + n, 1, n, variant=3)[0]  # shuffled, tie-free
    key = nodes["angle_z_q14"].astype(np.int64)
    valid = nodes["dist_mm_q2"] != 0
    bits = np.zeros(65536, np.uint8)
    bits[key[valid]] = 1
    assert bits.sum() == valid.sum()  # tie-free
    words = np.packbits(bits.reshape(2048, 32), axis=1, bitorder="little").view("<u4").reshape(2048)
    pop = np.array([bin(int(w)).count("1") for w in words])
    prefix = np.concatenate([[0], np.cumsum(pop)[:-1]])
    below = lambda k: (1 << (k & 31)) - 1
    rank = np.array([prefix[k >> 5] + bin(int(words[k >> 5]) & below(int(k))).count("1") for k in key[valid]])
    order = np.argsort(key[valid], kind="stable")
    assert (rank[order] == np.arange(valid.sum())).all()
    # a duplicate key makes the popcount fall short of the measured count
    if valid.sum() >= 2:
        k2 = key[valid].copy()
        k2[1] = k2[0]
        b2 = np.zeros(65536, np.uint8)
        b2[k2] = 1
        assert b2.sum() == valid.sum() - 1


# ---- round 2: shared final keys, Mode A as a scatter-min, CRC over byte ranges --------------------------------------
@pytest.mark.parametrize("seed", range(8))
def test_ranks_with_shared_keys_follow_the_stable_rule(seed):
    """scan_small.cu, ascended buffer: position of a node = (distinct keys below its key, from the bitmap)
    + (listed keys below its key: one entry per node beyond the first of a key) + (nodes with the same key earlier in
    the buffer) -- equals the stable sort."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 400))
    keys = rng.choice(65536, size=n, replace=False)
    for _ in range(int(rng.integers(0, 6))):  # a few shared keys, some shared by three
        i, j = rng.integers(0, n, 2)
        keys[j] = keys[i]
    distinct = np.unique(keys)
    listed = []  # what the rare-path sweep collects: every node beyond the first (in ANY order) of its key
    seen = set()
    for i in rng.permutation(n):
        if int(keys[i]) in seen:
            listed.append(int(keys[i]))
        seen.add(int(keys[i]))
    assert len(listed) == n - len(distinct)
    pos = np.empty(n, np.int64)
    for i in range(n):
        k = int(keys[i])
        pos[i] = np.searchsorted(distinct, k) + sum(1 for d in listed if d < k) + int((keys[:i] == k).sum())
    assert (pos[np.argsort(keys, kind="stable")] == np.arange(n)).all()


def mode_a_scatter_min_model(keys, dm_bits, quality, bins, m):
    """What the shared-memory Mode A kernel computes (two atomicMin sweeps), plus its conflict flag."""
    minv = np.full(m, 0xFFFFFFFF, np.uint64)
    for b, d in zip(bins, dm_bits):
        minv[b] = min(minv[b], d)
    wkey = np.full(m, 0xFFFFFFFF, np.uint64)
    conflict = False
    for k, d, q, b in zip(keys, dm_bits, quality, bins):
        if d == minv[b]:
            v = (int(k) << 8) | int(q)
            old = int(wkey[b])
            wkey[b] = min(old, v)
            conflict = conflict or ((old >> 8) == int(k) and old != v)
    return minv, wkey, conflict


@pytest.mark.parametrize("seed", range(10))
def test_mode_a_scatter_min_equals_the_sorted_walk(oracle, seed):
    """Mode A (reference rplidar_node.cpp:630-660) walks the measured points in ascending key order and keeps, per
    bin, the first point with the smallest dist_m.  Without any ordering: per bin min(dist_m bits), then min(key) among
    the points that hold it.  Equal keys only matter when two of them hold a bin's minimum with different qualities:
    that is the one case the kernel hands to the general kernel (conflict flag)."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(3, 500))
    keys = rng.integers(0, 65536, n) if seed % 2 else rng.choice(65536, size=n, replace=False)
    dist = rng.integers(1, 40, n) * 400  # few distinct distances: many ties in dist_m
    quality = rng.integers(0, 256, n)
    nodes = oracle.make_nodes(keys, dist, quality, 2)
    inv = seed % 3 == 0
    prm = oracle.scan_params(1, 1, int(inv), 0, 40.0, 0.1)
    order = np.argsort(keys, kind="stable")
    hdr, ranges, inten = oracle.publish(nodes[order], prm, stable=True)  # the serial walk over the sorted buffer
    m = hdr.beam_count
    assert m == n
    # bins exactly as the reference computes them (float chain), from the device-math proof helpers
    from test_device_math_proofs import exact_bins

    bins_of_key = exact_bins(m, inv)
    bins = bins_of_key[keys]
    dm_bits = (dist.astype(np.float32) / np.float32(4000.0)).view(np.uint32).astype(np.uint64)
    minv, wkey, conflict = mode_a_scatter_min_model(keys, dm_bits, quality, bins, m)
    dup_conflict = False  # the definition of the conflict: same key, same (minimal) dist_m, different quality
    for b in range(m):
        idx = [i for i in range(n) if bins[i] == b and dm_bits[i] == minv[b]]
        for x in idx:
            for y in idx:
                if keys[x] == keys[y] and quality[x] != quality[y]:
                    dup_conflict = True
    assert dup_conflict or not conflict  # the flag is only ever raised for a real one ...
    if not conflict:  # ... and without it the result is the serial walk's, even if an unflagged pair exists (a
        hit = minv != 0xFFFFFFFF  # smaller key of the same bin and distance took the bin from both)
        got_r = np.full(m, np.inf, np.float32)
        got_r[hit] = minv[hit].astype(np.uint32).view(np.float32)
        got_i = np.zeros(m, np.float32)
        got_i[hit] = (wkey[hit] & 0xFF).astype(np.float32)  # new protocol: intensity = quality
        assert (got_r.view(np.uint32) == ranges[:m].view(np.uint32)).all()
        assert (got_i.view(np.uint32) == inten[:m].view(np.uint32)).all()


def test_crc32_over_four_byte_ranges_combines_to_the_whole():
    """decode_formats.cu decode_hq_kernel: raw(s, A || B) = advance_|B|(raw(s, A)) ^ raw(0, B) for the table-driven
    reflected CRC-32; four ranges (192, 192, 192, 204 bytes incl. the SDK's zero padding) against zlib."""
    import zlib

    t0 = np.zeros(256, np.uint64)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (0xEDB88320 ^ (c >> 1)) if (c & 1) else (c >> 1)
        t0[i] = c

    def raw(s, data):
        for x in data:
            s = (s >> 8) ^ int(t0[(s ^ x) & 0xFF])
        return s

    def advance(s, nbytes):
        for _ in range(nbytes):
            s = (s >> 8) ^ int(t0[s & 0xFF])
        return s

    rng = np.random.default_rng(9)
    for _ in range(20):
        msg = bytes(rng.integers(0, 256, 777).astype(np.uint8)) + b"\0\0\0"
        r = [raw(0xFFFFFFFF, msg[0:192]), raw(0, msg[192:384]), raw(0, msg[384:576]), raw(0, msg[576:780])]
        crc = advance(advance(advance(r[0], 192) ^ r[1], 192) ^ r[2], 204) ^ r[3]
        assert (crc ^ 0xFFFFFFFF) == zlib.crc32(msg)
    # "advance" is linear, so it is four byte-indexed look-ups (what the kernel's tables hold)
    s = int(rng.integers(0, 1 << 32))
    parts = [advance(((s >> (8 * k)) & 0xFF) << (8 * k), 192) for k in range(4)]
    assert advance(s, 192) == parts[0] ^ parts[1] ^ parts[2] ^ parts[3]
