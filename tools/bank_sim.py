"""Offline shared-memory bank model of k_assemble_tiles (P1 triangle, elastic, C = 2).

Rebuilds the plan's tiles / records / work items for a structured criss-cross mesh on the host
(same rules as plan.cu: greedy tiles by incidence budget, records = distinct elements in ascending
id, items ordered by class, node degree, slot-in-row, node; record plane stride 0 mod 8) and counts LSU
wavefronts of the three conflicted
access streams of the kernel: the corner-coordinate gather of phase 1, the gradient gathers of
phase 2 and the image stores.  Used to evaluate plan-time layouts without GPU time:

    python tools/bank_sim.py [nx ny] [--color_nodes] [--color_slots]

prints the conflict factor (wavefronts / conflict-free wavefronts) of each stream; with the shipped layout
on a 401 x 101 mesh: corner gathers 1.68x (1.00x with --color_nodes, the plan's default), gradient gathers
1.19x, image stores 1.78x (ncu on config 2: 1.0x, 1.2x, 2.1-2.3x).

Bank model: a 128-bit access is served a quarter-warp (8 lanes) at a time; a quarter costs
max over the 8 16-byte bank groups of the number of DISTINCT addresses falling in it.
"""
import sys
import numpy as np

sys.path.insert(0, '/root/repo')
from fes_b200.mesh import rect_arrays  # noqa: E402


def rec_slot(r):
    return r ^ ((r >> 3) & 7)


def quarter_cost(addr16, active=None):
    """addr16: (n_lanes,) 16-byte chunk indices of one warp instruction (len multiple of 8, padded
    with -1 for inactive lanes).  Returns wavefronts."""
    total = 0
    for q in range(0, len(addr16), 8):
        a = addr16[q:q + 8]
        a = a[a >= 0]
        if len(a) == 0:
            continue
        u = np.unique(a)
        total += np.bincount(u % 8, minlength=8).max()
    return total


def build(nx, ny, inc_max=336):
    verts, el = rect_arrays([[0, 0], [4, 1]], (nx, ny))
    n_nodes, n_el = len(verts), len(el)
    # node -> incident (element, local index), ascending element id
    inc = [[] for _ in range(n_nodes)]
    for e in range(n_el):
        for a in range(3):
            inc[el[e, a]].append((e, a))
    # node rows: sorted neighbour list (incl. itself)
    nbr = []
    for v in range(n_nodes):
        s = set()
        for e, _ in inc[v]:
            s.update(el[e].tolist())
        nbr.append(sorted(s))
    return verts, el, inc, nbr


def tiles_of(inc, nbr, inc_max):
    tiles, v0, cnt = [], 0, 0
    n = len(inc)
    start = 0
    acc = 0
    for v in range(n):
        k = len(inc[v])
        if acc + k > inc_max or v - start >= 255:
            tiles.append((start, v))
            start, acc = v, 0
        acc += k
    tiles.append((start, n))
    return tiles


def item_split(cnt, L):
    need = (cnt + L - 1) // L
    lg = 0
    while (1 << lg) < need and lg < 5:
        lg += 1
    m = (cnt + (1 << lg) - 1) >> lg
    return lg, m


class Tile:
    pass


def make_tile(el, inc, nbr, v0, v1, L=2):
    T = Tile()
    elems = sorted({e for v in range(v0, v1) for e, _ in inc[v]})
    T.elems = elems
    T.rec_of = {e: r for r, e in enumerate(elems)}
    nodes = sorted({int(n) for e in elems for n in el[e]})
    T.nodes = nodes
    T.li_of = {n: i for i, n in enumerate(nodes)}
    # pairs and their contributions
    T.rows = []
    p0 = 0
    pairs = []
    for v in range(v0, v1):
        deg = len(nbr[v])
        for sl, w in enumerate(nbr[v]):
            contrib = []
            for e, a in inc[v]:
                row = el[e]
                for b in range(3):
                    if row[b] == w:
                        contrib.append((e, a, b))
            pairs.append(dict(v=v, sl=sl, deg=deg, q=2 * p0 + sl, contrib=contrib))
        p0 += deg
    T.pairs = pairs
    T.n_pairs = p0
    # items in plan order: class (group size, trips) descending, node degree, slot-in-row, node
    classes = sorted({item_split(len(p['contrib']), L) for p in pairs}, reverse=True)
    max_deg = max(p['deg'] for p in pairs)
    degrees = sorted({p['deg'] for p in pairs})
    by = {}
    for p in pairs:
        by[(p['v'], p['sl'])] = p
    items = []
    for (lg, m) in classes:
      for deg in degrees:
        for sl in range(max_deg):
            for v in range(v0, v1):
                p = by.get((v, sl))
                if p is None or p['deg'] != deg or item_split(len(p['contrib']), L) != (lg, m):
                    continue
                for j in range(1 << lg):
                    cs = p['contrib'][j * m:(j + 1) * m]
                    items.append(dict(pair=p, writer=(j == 0), lg=lg, contrib=cs))
    T.items = items
    return T


def sim_tile(T, el, slot_of=None, li_of=None, threads=128, GS=None, item_order=None, L=2):
    """Returns dict of wavefront counts."""
    ne = len(T.elems)
    if slot_of is None:
        slot_of = {r: rec_slot(r) for r in range(ne)}
    if li_of is None:
        li_of = T.li_of
    if GS is None:
        GS = ((ne + 7) & ~7) + 8   # plane stride 0 (mod 8), the last slot is the zero record (plan.cu)
    out = dict(coord=0, coord_ideal=0, gather=0, gather_ideal=0, store=0, store_ideal=0, p1store=0)
    # phase 1
    blk = 0
    while blk * threads < ne:
        for w in range(threads // 32):
            lanes = np.arange(32) + 32 * w
            gi = blk * threads + ((lanes + blk * (threads // 2)) % threads)
            live = gi < ne
            if not live.any():
                continue
            for k in range(3):
                addr = np.full(32, -1)
                for l in range(32):
                    if live[l]:
                        addr[l] = li_of[int(el[T.elems[gi[l]], k])]
                out['coord'] += quarter_cost(addr)
                out['coord_ideal'] += int(np.ceil(live.reshape(4, 8).any(axis=1).sum()))
            for n in range(3):
                addr = np.full(32, -1)
                for l in range(32):
                    if live[l]:
                        addr[l] = n * GS + slot_of[int(gi[l])]
                out['p1store'] += quarter_cost(addr)
        blk += 1
    # phase 2
    items = T.items if item_order is None else [T.items[i] for i in item_order]
    n_items = len(items)
    for t0 in range(0, n_items, 32):
        warp = items[t0:t0 + 32]
        same = all(a == b for it in warp for (_, a, b) in it['contrib'])
        for x in range(L):
            if not any(len(it['contrib']) > x for it in warp):
                continue
            aa = np.full(32, -1)
            bb = np.full(32, -1)
            for l, it in enumerate(warp):
                if len(it['contrib']) > x:
                    e, a, b = it['contrib'][x]
                    s = slot_of[T.rec_of[e]]
                    aa[l] = a * GS + s
                    bb[l] = b * GS + s
                else:
                    aa[l] = bb[l] = GS - 1
            out['gather'] += quarter_cost(aa)
            out['gather_ideal'] += 4 if len(warp) == 32 else (len(warp) + 7) // 8
            if not same:
                out['gather'] += quarter_cost(bb)
                out['gather_ideal'] += 4 if len(warp) == 32 else (len(warp) + 7) // 8
        # image stores (writer lanes): two STS.128
        r0 = np.full(32, -1)
        r1 = np.full(32, -1)
        for l, it in enumerate(warp):
            if it['writer']:
                tid = (t0 + l) % threads
                swap = (tid >> 2) & 1
                q0, deg = it['pair']['q'], it['pair']['deg']
                r0[l] = q0 + (deg if swap else 0)
                r1[l] = q0 + (0 if swap else deg)
        for r in (r0, r1):
            out['store'] += quarter_cost(r)
            out['store_ideal'] += int(((r.reshape(4, 8) >= 0).any(axis=1)).sum())
    return out


def color_nodes(T, el, threads=128):
    """Greedy: local index of the tile's distinct nodes so that the 8 lanes of every quarter-warp
    of phase 1 read 8 different bank groups (or the same address)."""
    ne = len(T.elems)
    groups = []
    blk = 0
    while blk * threads < ne:
        for w in range(threads // 32):
            lanes = np.arange(32) + 32 * w
            gi = blk * threads + ((lanes + blk * (threads // 2)) % threads)
            for q in range(4):
                g = [int(x) for x in gi[8 * q:8 * q + 8] if x < ne]
                for k in range(3):
                    groups.append({int(el[T.elems[r], k]) for r in g})
        blk += 1
    node_groups = {n: [] for n in T.nodes}
    for gi_, g in enumerate(groups):
        for n in g:
            node_groups[n].append(gi_)
    color, used = {}, [np.zeros(8, int) for _ in groups]
    fill = np.zeros(8, int)
    cap = 32  # 8 colours x 32 = 256 local indices
    for n in T.nodes:
        cost = np.zeros(8)
        for gi_ in node_groups[n]:
            cost += used[gi_]
        cost = cost + 1e-3 * fill + 1e6 * (fill >= cap)
        c = int(np.argmin(cost))
        color[n] = c
        fill[c] += 1
        for gi_ in node_groups[n]:
            used[gi_][c] += 1
    nxt = np.zeros(8, int)
    li = {}
    for n in T.nodes:
        c = color[n]
        li[n] = c + 8 * nxt[c]
        nxt[c] += 1
    return li


def gather_groups(T, items, L=2):
    """(plane shift, record) sets read together by a quarter-warp in phase 2."""
    groups = []
    for t0 in range(0, len(items), 32):
        warp = items[t0:t0 + 32]
        same = all(a == b for it in warp for (_, a, b) in it['contrib'])
        for q in range(0, len(warp), 8):
            lanes = warp[q:q + 8]
            for x in range(L):
                ga, gb = set(), set()
                for it in lanes:
                    if len(it['contrib']) > x:
                        e, a, b = it['contrib'][x]
                        ga.add((a, T.rec_of[e]))
                        gb.add((b, T.rec_of[e]))
                if ga:
                    groups.append(ga)
                if gb and not same:
                    groups.append(gb)
    return groups


def color_slots(T, items, sweeps=3, L=2):
    """Greedy + refinement: slot (mod 8) of every record so quarter-warp gathers spread over banks."""
    ne = len(T.elems)
    groups = gather_groups(T, items, L)
    rec_groups = [[] for _ in range(ne)]
    for gi_, g in enumerate(groups):
        for (a, r) in g:
            rec_groups[r].append((gi_, a))
    used = np.zeros((len(groups), 8), int)
    col = np.full(ne, -1)
    fill = np.zeros(8, int)
    cap = (ne + 7) // 8

    def cost_of(r, c):
        s = 0
        for gi_, a in rec_groups[r]:
            s += used[gi_, (a + c) % 8]
        return s

    order = np.argsort([-len(g) for g in rec_groups], kind='stable')
    for r in order:
        best, bc = None, 0
        for c in range(8):
            if fill[c] >= cap:
                continue
            k = cost_of(r, c) + 1e-3 * fill[c]
            if best is None or k < best:
                best, bc = k, c
        col[r] = bc
        fill[bc] += 1
        for gi_, a in rec_groups[r]:
            used[gi_, (a + bc) % 8] += 1
    for _ in range(sweeps):
        for r in range(ne):
            c0 = col[r]
            for gi_, a in rec_groups[r]:
                used[gi_, (a + c0) % 8] -= 1
            fill[c0] -= 1
            best, bc = None, c0
            for c in range(8):
                if fill[c] >= cap:
                    continue
                k = cost_of(r, c)
                if best is None or k < best or (k == best and c == c0):
                    best, bc = k, c
            col[r] = bc
            fill[bc] += 1
            for gi_, a in rec_groups[r]:
                used[gi_, (a + bc) % 8] += 1
    nxt = np.zeros(8, int)
    slot = {}
    for r in range(ne):
        c = col[r]
        slot[r] = int(c + 8 * nxt[c])
        nxt[c] += 1
    return slot


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    nx, ny = (int(args[0]), int(args[1])) if len(args) >= 2 else (401, 101)
    policy = [a[2:] for a in sys.argv[1:] if a.startswith('--')]
    verts, el, inc, nbr = build(nx, ny)
    tl = tiles_of(inc, nbr, 336)
    tot = {}
    n_sim = 0
    for (v0, v1) in tl[len(tl) // 4: len(tl) // 4 + 40]:
        T = make_tile(el, inc, nbr, v0, v1)
        li = color_nodes(T, el) if 'color_nodes' in policy else None
        slot = color_slots(T, T.items) if 'color_slots' in policy else None
        GS = None
        if slot is not None:
            GS = ((max(slot.values()) + 1 + 7) & ~7) + 1
        r = sim_tile(T, el, slot_of=slot, li_of=li, GS=GS)
        for k, v in r.items():
            tot[k] = tot.get(k, 0) + v
        n_sim += 1
    print('tiles', len(tl), 'simulated', n_sim, 'records/tile', len(T.elems), 'items/tile', len(T.items), 'nodes', len(T.nodes))
    for k in ('coord', 'gather', 'store'):
        print(f'{k:8s} wavefronts {tot[k]:8d} ideal {tot[k + "_ideal"]:8d} factor {tot[k] / tot[k + "_ideal"]:.3f}')
    print('p1store', tot['p1store'])


if __name__ == '__main__':
    main()
