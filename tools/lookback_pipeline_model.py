#!/usr/bin/env python
"""Discrete-event model of local_sort_fine_kernel's persistent two-tile pipeline and its three-level tile
prefix (CPU only; no GPU needed).  296 CTAs draw tickets; per tile: fix-up .2P, publish, front of the
next tile .3P, resolve (needs every earlier tile of the group published, the earlier groups' sums and
the previous super-group's prefix; costs L, one look-back round trip), emission .3P, placement .2P.
A fraction p2 of the tickets are tiles without output (published at the draw).  `arrival=True` models
group sums published by the last ARRIVING tile instead of by the group's last tile's resolve.

What it shows (profiles/r02_local_sort_one_bucket_tiles.md): the pipeline degrades gently while L is a
few percent of P (config 2: P ~ 17-23 us per tile) and super-linearly once L reaches 0.2-0.4 P -- the
one-bucket tiles of config 3 on 8 GPUs have P ~ 9-14 us, and the measured look-back polls take 1.5-4 us.
usage: lookback_pipeline_model.py"""
import heapq, random, sys
def simulate(n_tiles, p2, P, n_cta=296, L=0.7, jitter=0.1, seed=1, group=128, sup=32, pend_resolve=True, arrival=False):
    """Discrete-event model of local_sort_fine_kernel's pipeline; times in microseconds.
    P: work per tile; phases fixup .2P, front .3P, emit .3P, place .2P; L = one L2 round trip."""
    rnd = random.Random(seed)
    kind2 = [rnd.random() < p2 for _ in range(n_tiles)]
    pub = [None]*n_tiles      # publish time
    gsum = {}                 # group -> time available
    ticket = 0
    INF = float('inf')
    # state machine per CTA; we process CTAs in global time order using a heap of (time, cta, phase)
    # phases: 0 draw N, 1 after fixup -> publish A, then front N, 2 resolve A (may block), 3 emit+place -> loop
    A = [None]*n_cta; N = [None]*n_cta; pend=[[] for _ in range(n_cta)]
    heap = []
    waiting = []  # (cta) blocked in resolve
    def dur(x): return x*P*(1+jitter*(rnd.random()*2-1))
    def draw(t, c):
        nonlocal ticket
        lst=[]
        while True:
            if ticket >= n_tiles: return None, t
            x = ticket; ticket += 1; t += 0.3  # ticket + bounds loads
            if kind2[x]:
                pub[x] = t
                if x % group == group-1: lst.append(x)
                continue
            pend[c] = lst
            return x, t
    def group_ready_time(g_tile_last):
        pass
    # init
    for c in range(n_cta):
        a, t = draw(0.0, c)
        A[c] = a
        if a is None: continue
        t += dur(0.3) + dur(0.2)  # front + placement
        heapq.heappush(heap, (t, c, 0))
    done = 0; wait_total = 0.0; tmax = 0
    def arrival_gsum(gg):
        lo, hi = gg*group, min(n_tiles, (gg+1)*group)
        t = 0.0
        for b in range(lo, hi):
            if pub[b] is None: return None
            t = max(t, pub[b])
        return t + L
    def deps_time(a):
        """time at which everything tile a's resolve needs is available, or None if unknown yet"""
        g = a // group
        t = 0.0
        if arrival:
            for gg in range((g // sup) * sup, g):
                if gg not in gsum:
                    v = arrival_gsum(gg)
                    if v is None: return None
                    gsum[gg] = v
        for b in range(g*group, a):
            if pub[b] is None: return None
            t = max(t, pub[b])
        # group sums of earlier groups in the super-group
        s0 = (g // sup) * sup
        for gg in range(s0, g):
            if gg not in gsum: return None
            t = max(t, gsum[gg])
        if s0 > 0:
            if ('s', s0//sup - 1) not in gsum: return None
            t = max(t, gsum[('s', s0//sup-1)])
        return t
    def try_group_sum(a, tnow):
        """last tile of group: its resolve publishes gsum once all group tiles are published"""
        g = a // group
        t = tnow
        for b in range(g*group, a):
            if pub[b] is None: return None
            t = max(t, pub[b])
        return t + L
    blocked = []  # (cta, t_arrive, tile, is_pend)
    def attempt(c, t, a):
        """returns finish time of resolve of tile a started at t, or None if blocked"""
        g = a // group
        if not arrival and a % group == group-1 and g not in gsum:
            tg = try_group_sum(a, t)
            if tg is None: return None
            gsum[g] = tg
        d = deps_time(a)
        if d is None: return None
        fin = max(t, d) + L
        if a % group == group-1 and (g % sup) == sup-1:
            gsum[('s', g//sup)] = fin
        return fin
    progress = True
    while heap or blocked:
        if not heap:
            # retry blocked
            nb=[]; moved=False
            for (c,t,a,ph) in blocked:
                fin = attempt(c, t, a)
                if fin is None: nb.append((c,t,a,ph))
                else:
                    heapq.heappush(heap,(fin,c,ph)); moved=True
            blocked = nb
            if not moved: raise RuntimeError("deadlock %d blocked" % len(blocked))
            continue
        t, c, ph = heapq.heappop(heap)
        tmax = max(tmax, t)
        if ph == 0:
            n, t = draw(t, c); N[c] = n
            t += dur(0.2)          # fixup(A)
            pub[A[c]] = t           # publish(A)
            if n is not None: t += dur(0.3)   # front(N)
            heapq.heappush(heap, (t, c, 2))
            # a publish may unblock others
            nb=[]
            for (c2,t2,a2,ph2) in blocked:
                fin = attempt(c2, t2, a2)
                if fin is None: nb.append((c2,t2,a2,ph2))
                else: heapq.heappush(heap,(fin,c2,ph2))
            blocked = nb
        elif ph == 2:   # resolve A
            fin = attempt(c, t, A[c])
            if fin is None:
                blocked.append((c,t,A[c],2)); continue
            if fin > t:  # re-queue at fin to keep causality
                wait_total += fin - t - L
                heapq.heappush(heap,(fin,c,3)); 
            else:
                heapq.heappush(heap,(fin,c,3))
        elif ph == 3:   # emission, pend resolves, placement
            t += dur(0.3)
            ok=True
            for x in pend[c]:
                fin = attempt(c, t, x)
                if fin is None:
                    blocked.append((c,t,x,4)); ok=False; break
                t = fin
            if not ok: continue
            pend[c]=[]
            done += 1
            if N[c] is None: continue
            t += dur(0.2)
            A[c] = N[c]
            heapq.heappush(heap,(t,c,0))
            nb=[]
            for (c2,t2,a2,ph2) in blocked:
                fin = attempt(c2, t2, a2)
                if fin is None: nb.append((c2,t2,a2,ph2))
                else: heapq.heappush(heap,(fin,c2,ph2))
            blocked = nb
        elif ph == 4:  # blocked pend resolved; continue phase 3 remainder
            pend[c]=pend[c][1:]
            heapq.heappush(heap,(t,c,3)) if False else None
            # simplified: finish iteration
            done += 1
            if N[c] is None: continue
            t += dur(0.2); A[c]=N[c]; heapq.heappush(heap,(t,c,0))
    return tmax, wait_total

if __name__ == "__main__":
    for name, nt, p2, P in [("cfg2 multi-bucket", 16400, 0.0, 17.0), ("one-bucket 240M", 73000, 0.10, 9.0), ("one-bucket 387M", 71000, 0.08, 14.0)]:
        for L in (0.7, 2.0, 4.0):
            for arr in (False, True):
                tm, w = simulate(nt, p2, P, L=L, arrival=arr)
                ideal = nt*(1-p2)*P/296
                print(f"{name:22s} L={L} arrival={arr}: total {tm/1e3:7.2f} ms  ideal {ideal/1e3:6.2f} ms  ratio {tm/ideal:5.2f}")
