"""CPU restatement of the device-side index logic that decides WHICH tile a CTA works on — a wrong formula there silently
drops or duplicates tiles.  Mirrors (line for line, in Python integers):
  * decode_item / make_tile_sched   (flash_fwd_common.cuh): the persistent forward's static work list
  * item_geom's per-tile key-block counts vs the element-wise causal rule j - i <= sk - sq (mask.h:20-72 in the reference)
  * the fused backward's rotated walk over the 64-row query sub-tiles (flash_bwd_tc_sm100.cu: it0, steps_per_head, rot, q_sub)
  * the forward's role protocol in COUNTS (flash_fwd_p4_sm100.cu): tokens given / taken by the two MMA warps, K/V ring releases
    per tile and warp, Q loads vs waits, staged tiles the store warp expects vs the softmax slots hand over — a mismatch there
    is a hang, not a wrong number
"""
import itertools
import random


def make_tile_sched(b, h, sq, sk, d, group_mb=16, causal=False):
    num_mblk = (sq + 255) // 256
    bh = b * h
    kv_bytes_per_head = 2 * sk * d * 2
    grp = (group_mb << 20) // kv_bytes_per_head if kv_bytes_per_head > 0 else bh
    grp = max(1, min(grp, bh))
    levels = (num_mblk + 1) // 2 if causal else num_mblk          # causal: a slot pairs row blocks L and num_mblk - 1 - L
    return {"num_mblk": num_mblk, "levels": levels, "paired": causal, "bh": bh, "group": grp, "total": levels * bh}


def decode_item(ts, n, h, causal):
    half, slot = (n & 1, n >> 1) if ts["paired"] else (0, n)
    per_group = ts["group"] * ts["levels"]
    g = slot // per_group
    r = slot - g * per_group
    heads_here = min(ts["group"], ts["bh"] - g * ts["group"])
    level = r // heads_here
    head_local = r - level * heads_here
    bhi = g * ts["group"] + head_local
    if ts["paired"]:
        heavy = ts["num_mblk"] - 1 - level
        mblk = heavy if half == 0 else (level if level < heavy else ts["num_mblk"])   # num_mblk = void item
    else:
        mblk = level
    return mblk, bhi % h, bhi // h


def cta_items(ts, cta, grid):
    """the static schedule of one CTA (flash_fwd_p4_sm100.cu: static_item)"""
    slots = range(cta, ts["total"], grid)
    return [2 * s + half for s in slots for half in (0, 1)] if ts["paired"] else list(slots)


def test_forward_work_list_covers_every_tile_pair_exactly_once():
    random.seed(0)
    shapes = [(4, 32, 4096, 4096, 128), (4, 32, 8192, 8192, 128), (1, 1, 1, 1, 128), (3, 6, 129, 127, 64),
              (2, 4, 1025, 1, 128), (256, 32, 16384, 16384, 128), (1, 2, 768, 768, 128)]
    shapes += [(random.randint(1, 9), random.randint(1, 33), random.randint(1, 5000), random.randint(1, 70000),
                random.choice([64, 128])) for _ in range(200)]
    for (b, h, sq, sk, d), causal, mb in itertools.product(shapes, (False, True), (1, 16, 4096)):
        if b * h * ((sq + 255) // 256) > 300000:
            b = 2                                   # keep the enumeration small; the formulas do not depend on b's size
        ts = make_tile_sched(b, h, sq, sk, d, mb, causal)
        seen = set()
        grid = min(ts["total"], 148)
        for cta in range(grid):
            for n in cta_items(ts, cta, grid):
                item = decode_item(ts, n, h, causal)
                mblk, bidh, bidb = item
                if mblk == ts["num_mblk"]:
                    continue                        # void second half of an unpaired middle block
                assert 0 <= mblk < ts["num_mblk"] and 0 <= bidh < h and 0 <= bidb < b, (item, ts)
                assert item not in seen, f"duplicate {item} for {(b, h, sq, sk, d, causal, mb)}"
                seen.add(item)
        assert len(seen) == ts["num_mblk"] * b * h


def test_causal_schedule_is_balanced():
    """paired causal slots all cost 2 * num_mblk + 2 key steps (sq == sk): per-CTA sums differ by at most one slot"""
    for (b, h, s) in [(4, 32, 8192), (4, 16, 16384), (1, 7, 4096), (2, 3, 2048 + 256)]:
        ts = make_tile_sched(b, h, s, s, 128, 16, True)
        grid = min(ts["total"], 148)
        cost = []
        for cta in range(grid):
            c = 0
            for n in cta_items(ts, cta, grid):
                mblk, _, _ = decode_item(ts, n, h, True)
                if mblk < ts["num_mblk"]:
                    c += nblk(mblk * 256, s, s, True) + nblk(mblk * 256 + 128, s, s, True)
            cost.append(c)
        per_slot = 4 * ts["num_mblk"] + 2 if ts["num_mblk"] % 2 == 0 else None
        if per_slot:
            assert max(cost) - min(cost) <= per_slot, (b, h, s, min(cost), max(cost))


def nblk(mt, sq_b, sk_b, causal):
    """item_geom: key blocks a 128-row query tile starting at row mt has to visit"""
    kv_end = sk_b if mt < sq_b else 0
    if causal:
        kv_end = min(kv_end, max(0, mt + 128 + (sk_b - sq_b)))
    return (kv_end + 127) // 128


def test_key_block_counts_agree_with_the_elementwise_causal_rule():
    random.seed(1)
    cases = [(1, 1), (63, 65), (65, 63), (128, 128), (129, 127), (1023, 1025), (1025, 1023), (1, 1025), (1025, 1), (300, 1024)]
    cases += [(random.randint(1, 1500), random.randint(1, 1500)) for _ in range(150)]
    for (sq, sk), causal in itertools.product(cases, (False, True)):
        off = sk - sq
        for mt in range(0, sq, 128):
            rows = range(mt, min(mt + 128, sq))
            last_visible = max((min(sk - 1, i + off) if causal else sk - 1) for i in rows)   # -1 or less: nothing visible
            need = 0 if last_visible < 0 else last_visible // 128 + 1
            assert nblk(mt, sq, sk, causal) == need, (sq, sk, causal, mt)


def fused_bwd_steps(n0, sq_b, sk_b, causal, block_x):
    """query sub-tiles (64 rows) the fused backward's CTA for key tile n0 walks, in its rotated order"""
    off = sk_b - sq_b
    i_first = max(0, n0 - off) if causal else 0
    it0 = i_first // 64
    nsub = (sq_b + 63) // 64
    steps = max(0, nsub - it0)
    rot = (block_x * 2) % steps if steps > 0 else 0
    order = []
    for st in range(steps):
        ls = st % steps + rot
        if ls >= steps:
            ls -= steps
        order.append(it0 + ls)
    return order


def test_fused_backward_walk_visits_every_visible_query_sub_tile_once():
    random.seed(2)
    cases = [(4096, 4096), (300, 1024), (1025, 1023), (1, 1), (65, 63), (129, 127), (1, 1025), (1025, 1)]
    cases += [(random.randint(1, 3000), random.randint(1, 3000)) for _ in range(100)]
    for (sq, sk), causal in itertools.product(cases, (False, True)):
        off = sk - sq
        for bx, n0 in enumerate(range(0, sk, 128)):
            order = fused_bwd_steps(n0, sq, sk, causal, bx)
            assert len(order) == len(set(order))
            # every sub-tile that holds a query row able to see a key of this tile must be walked
            keys = range(n0, min(n0 + 128, sk))
            for it in range((sq + 63) // 64):
                rows = range(it * 64, min(it * 64 + 64, sq))
                visible = any((not causal) or (keys[0] <= i + off) for i in rows)
                assert (it in order) or not visible, (sq, sk, causal, n0, it)


def test_backward_causal_predicates_agree_with_the_elementwise_rule():
    """flash_bwd_tc_sm100.cu, dK/dV-type kernels: thread (g, r) owns key row jg = n0 + r and query columns 32g..32g+31 of the
    64-row sub-tile `it`; it zeroes P where `need_mask && column < cmin`.  Truth: key j is visible to query i iff j <= i + off."""
    random.seed(3)
    cases = [(300, 1024), (1025, 1023), (65, 63), (129, 127), (128, 128), (1, 1025), (1025, 1), (640, 640)]
    cases += [(random.randint(1, 700), random.randint(1, 700)) for _ in range(40)]
    for sq, sk in cases:
        off = sk - sq
        for bx, n0 in enumerate(range(0, sk, 128)):
            for it in fused_bwd_steps(n0, sq, sk, True, bx):
                need_mask = it * 64 < n0 + 128 - 1 - off
                for r in (0, 1, 31, 63, 64, 127):
                    jg = n0 + r
                    if jg >= sk:
                        continue
                    for g in (0, 1):
                        cmin = (jg - off - it * 64 - g * 32) if need_mask else 0
                        for c in range(32):
                            i = it * 64 + g * 32 + c
                            if i >= sq:
                                continue
                            kernel_zeroes = need_mask and c < cmin
                            assert kernel_zeroes == (jg > i + off), (sq, sk, n0, it, r, g, c)
            # sub-tiles the walk skips hold no query row that sees a key of this tile
            walked = set(fused_bwd_steps(n0, sq, sk, True, bx))
            for it in range((sq + 63) // 64):
                if it not in walked:
                    assert all(n0 > i + off for i in range(it * 64, min(it * 64 + 64, sq)))


def test_dq_kernel_key_ranges_and_mask_agree_with_the_elementwise_rule():
    """flash_bwd_dq_kernel_sm100 / forward: a 128-row query tile at m0 visits ceil(min(sk, m0 + 128 + off) / 128) key tiles and
    masks column c of key tile n0 for row `row` iff c > min(sk - 1, row + off) - n0 (only on tiles flagged need_mask)."""
    random.seed(4)
    cases = [(300, 1024), (1025, 1023), (65, 63), (129, 127), (1, 1025), (1025, 1)] + [(random.randint(1, 600), random.randint(1, 600)) for _ in range(40)]
    for (sq, sk), causal in itertools.product(cases, (False, True)):
        off = sk - sq
        for m0 in range(0, sq, 128):
            nb = nblk(m0, sq, sk, causal)
            for row in {m0, min(m0 + 127, sq - 1), min(m0 + 64, sq - 1)}:
                col_limit = min(sk - 1, row + off) if causal else sk - 1
                for j in range(nb):
                    n0 = j * 128
                    need_mask = (n0 + 128 > sk) or (causal and (n0 + 128 - 1 > m0 + off))
                    for c in (0, 1, 63, 64, 127):
                        key = n0 + c
                        truth_visible = key < sk and ((not causal) or key <= row + off)
                        kernel_visible = not (need_mask and c > col_limit - n0)
                        assert kernel_visible == truth_visible, (sq, sk, causal, m0, row, j, c)
                # nothing visible beyond the visited tiles
                assert all(not (k < sk and ((not causal) or k <= row + off)) for k in range(nb * 128, min(sk, nb * 128 + 256)))


def _make_fastdiv(d):
    """restatement of make_fastdiv (flash_fwd_common.cuh): divisor -> (mul, shr) with n // d == (n * mul >> 32) >> shr"""
    if d == 1:
        return 0, 0
    lg = 0
    while (1 << lg) < d:
        lg += 1
    pw = 31 + lg
    return ((1 << pw) + d - 1) // d, pw - 32


def test_fastdiv_matches_integer_division_below_2_31():
    """the work-list decode divides by (group * num_mblk), the group size and the head count with multiply-high + shift"""
    import random
    rnd = random.Random(0)
    divisors = list(range(1, 300)) + [2 ** k for k in range(1, 31)] + [2 ** k - 1 for k in range(2, 31)] + [2 ** k + 1 for k in range(1, 30)] \
        + [rnd.randrange(1, 2 ** 31) for _ in range(300)]
    for d in divisors:
        mul, shr = _make_fastdiv(d)
        assert mul < 2 ** 32, d
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1, 2 ** 31 - d, (2 ** 31 - 1) // d * d, (2 ** 31 - 1) // d * d - 1] \
            + [rnd.randrange(0, 2 ** 31) for _ in range(200)]
        for n in ns:
            if 0 <= n < 2 ** 31:
                q = n if d == 1 else ((n * mul) >> 32) >> shr
                assert q == n // d, (n, d, q, n // d)


# ------------------------------------------------------------------------------------------------------------------
# The forward's role protocol in counts (flash_fwd_p4_sm100.cu): every role walks the same item list and derives, from the
# item's geometry alone, how many times it will wait on / arrive at each barrier.  A mismatch is a hang, not a wrong number
# (the retry-list hang of round 2 was exactly that: roles that disagreed about the list).  Restated here per item:
#   tokens     (0,j) takes iff j >= 1 and j-1 < nb1, gives iff j < nb1;  (1,j) takes iff j < nb0, gives iff j+1 < nb0
#   K/V ring   every tile of the item (2 * max(nb0, nb1)) is released exactly once by EACH MMA warp (commit or plain arrive)
#   Q buffers  the producer loads Q_t iff nb_t > 0; MMA warp t waits for it iff nb_t > 0
#   hand-over  the store warp expects one staged tile per (item, t) with rows and keys; the softmax slot t hands over the same
#   S / P      softmax slot t waits for nb_t score tiles and arrives 4 quarters each; MMA warp t issues nb_t S and waits 4 nb_t quarters
# ------------------------------------------------------------------------------------------------------------------
def _item_protocol_counts(nb0, nb1):
    nb = (nb0, nb1)
    takes, gives = [0, 0], [0, 0]
    for j in range(nb0):                                   # MMA warp of tile 0
        takes[0] += (j >= 1 and j - 1 < nb1)
        gives[0] += (j < nb1)                              # wakes (1, j)
    for j in range(nb1):                                   # MMA warp of tile 1
        takes[1] += (j < nb0)
        gives[1] += (j + 1 < nb0)                          # wakes (0, j + 1)
    nbmax = max(nb)
    releases = []
    for t in range(2):
        rel = [0] * (2 * nbmax)                            # ring tiles of the item: K_j = 2 j, V_j = 2 j + 1
        if nb[t] > 0:
            rel[0] += 1                                    # K_0 right after S(0)
        for j in range(nb[t]):
            rel[2 * j + 1] += 1                            # V_j after block j
            if j + 1 < nb[t]:
                rel[2 * j + 2] += 1                        # K_j+1 after S(j+1)
        for j in range(nb[t], nbmax):                      # tiles only the other query tile needs
            rel[2 * j] += 1
            rel[2 * j + 1] += 1
        releases.append(rel)
    return takes, gives, releases


def test_forward_role_protocol_counts_are_consistent_for_every_block_count_pair():
    for nb0, nb1 in itertools.product(range(0, 12), repeat=2):
        takes, gives, releases = _item_protocol_counts(nb0, nb1)
        assert gives[0] == takes[1], (nb0, nb1, "tokens passed to tile 1's MMA warp")
        assert gives[1] == takes[0], (nb0, nb1, "tokens passed to tile 0's MMA warp")
        for t in range(2):
            assert all(c == 1 for c in releases[t]), (nb0, nb1, t, releases[t])   # count-2 barrier: one arrival per warp and use


def test_every_role_derives_the_same_work_from_an_item():
    """producer, the two MMA warps, the softmax slots and the store warp all decode (item -> geometry) themselves; what they
    conclude has to agree for every item of every CTA, in pass 0 and — the same static list again — in the retry pass."""
    random.seed(1)
    shapes = [(5, 8, 2211, 1202, 64, True), (6, 8, 2070, 1777, 64, False), (4, 16, 2275, 1261, 64, True), (2, 4, 300, 333, 128, True),
              (1, 2, 1, 1, 64, False), (3, 4, 1025, 1, 128, True), (2, 2, 129, 4000, 128, False)]
    shapes += [(random.randint(1, 6), random.randint(1, 12), random.randint(1, 2600), random.randint(1, 2600), random.choice([64, 128]),
                random.random() < 0.5) for _ in range(150)]
    for b, h, sq, sk, d, causal in shapes:
        ts = make_tile_sched(b, h, sq, sk, d, 16, causal)
        grid = min(ts["total"], 148)
        for cta in range(grid):
            q_loads, q_waits, stores, handovers, s_issued, s_waited = [0, 0], [0, 0], 0, 0, [0, 0], [0, 0]
            kv_loaded = kv_released = 0
            for n in cta_items(ts, cta, grid):
                mblk, _, _ = decode_item(ts, n, h, causal)
                m0 = mblk * 256
                if m0 >= sq:                               # g.skip: every role skips the item
                    continue
                nb = [nblk(m0 + 128 * t, sq, sk, causal) for t in range(2)]
                n_blocks = max(nb)
                if n_blocks > 0:                           # producer and MMA warps (they skip items without keys)
                    kv_loaded += 2 * n_blocks
                    _, _, releases = _item_protocol_counts(nb[0], nb[1])
                    kv_released += sum(releases[0])
                    assert sum(releases[0]) == sum(releases[1]) == 2 * n_blocks
                for t in range(2):
                    mt = m0 + 128 * t
                    has_rows = mt < sq
                    if n_blocks > 0 and nb[t] > 0:
                        q_loads[t] += 1                    # producer: load_q(t)
                        q_waits[t] += 1                    # MMA warp t: bar_q_full
                        s_issued[t] += nb[t]
                    if has_rows and nb[t] > 0:             # softmax slot t: key loop + epilogue
                        s_waited[t] += nb[t]
                        handovers += 1
                    if not (mt >= sq or nb[t] == 0):       # store warp's count (dense: every such tile is a whole TMA tile)
                        stores += 1
                    assert not (nb[t] > 0 and not has_rows), "a tile without rows never has key blocks"
            assert q_loads == q_waits and s_issued == s_waited and stores == handovers and kv_loaded == kv_released, \
                (b, h, sq, sk, d, causal, cta)
