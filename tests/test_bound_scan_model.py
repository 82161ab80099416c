"""Host-side model of the two-ended bounded near/far scan (``csrc/grid.cu``: ``bound_dir_kernel<false>`` then
``bound_dir_kernel<true>``).

The CUDA threads of one ray (one per 32-sample segment and direction) run concurrently and prune each other through
the ``bnear`` / ``bfar`` words they update with atomicMin / atomicMax.  The claim in DESIGN.md section 4.1 is that a
STALE read of those words only costs work, never correctness: whatever the interleaving, the result equals
min / max of the depths of all samples with ``ds < thresh`` (``models/renderer.py:86-95``).  This test restates the
thread logic in Python and checks the claim against brute force under random schedules, random hit patterns and random
"certificate" codes (1 = proven miss, 2 = proven hit, 0 = evaluate), counting evaluations on the way."""
import random

import numpy as np
import pytest

SEG = 32
INF = float("inf")


class Ray:
    def __init__(self, hit, code, depth):
        self.hit, self.code, self.depth = hit, code, depth
        self.bnear, self.bfar = INF, -1.0      # ray_setup_kernel: +inf / -1
        self.evaluated = 0


def forward_thread(ray, g, read_lag):
    """Generator: one yield per sample step, so that the scheduler can interleave the threads of a ray."""
    n = len(ray.hit)
    s0, s1 = g * SEG, min((g + 1) * SEG, n)
    seen = ray.bnear                      # what this thread last read (may be stale)
    if seen < ray.depth[s0]:
        return
    for s in range(s0, s1):
        yield
        c = ray.code[s]
        if c == 1:
            continue
        hit = c == 2
        if not hit:
            if random.random() > read_lag:   # re-read the word (otherwise keep the stale value)
                seen = ray.bnear
            if seen < ray.depth[s]:
                return
            ray.evaluated += 1
            hit = ray.hit[s]
        if hit:
            ray.bnear = min(ray.bnear, ray.depth[s])
            return


def backward_thread(ray, g, read_lag):
    n = len(ray.hit)
    s0, s1 = g * SEG, min((g + 1) * SEG, n)
    first = ray.bnear                      # final: launch 2 starts after launch 1 has finished
    if first == INF or ray.depth[s1 - 1] < first:
        return
    seen = ray.bfar
    if seen >= ray.depth[s1 - 1]:
        return
    for s in range(s1 - 1, s0 - 1, -1):
        yield
        if ray.depth[s] <= first:
            ray.bfar = max(ray.bfar, first)
            return
        c = ray.code[s]
        if c == 1:
            continue
        hit = c == 2
        if not hit:
            if random.random() > read_lag:
                seen = ray.bfar
            if seen >= ray.depth[s]:
                return
            ray.evaluated += 1
            hit = ray.hit[s]
        if hit:
            ray.bfar = max(ray.bfar, ray.depth[s])
            return


def run_launch(threads):
    """Random interleaving of the threads' steps (any subset may be 'resident' at a time)."""
    live = list(threads)
    while live:
        t = random.choice(live)
        try:
            next(t)
        except StopIteration:
            live.remove(t)


def make_case(rng, n=256):
    kind = rng.integers(0, 6)
    hit = np.zeros(n, bool)
    if kind == 1:
        hit[rng.integers(0, n)] = True
    elif kind == 2:                      # one interval (a ray through a thin shell)
        a = rng.integers(0, n)
        hit[a:min(n, a + rng.integers(1, 60))] = True
    elif kind == 3:                      # two intervals (entry shell, core, exit shell)
        a, b = sorted(rng.integers(0, n, 2))
        hit[a:a + rng.integers(1, 40)] = True
        hit[b:b + rng.integers(1, 40)] = True
    elif kind == 4:
        hit = rng.random(n) < 0.05
    elif kind == 5:
        hit[:] = True
    # certificate codes: only where they are TRUE statements about the sample (the grid is sound), random coverage
    known = rng.random(n) < rng.choice([0.0, 0.5, 0.9, 1.0])
    code = np.where(known, np.where(hit, 2, 1), 0)
    depth = np.cumsum(rng.random(n).astype(np.float32) * 0.01 + 1e-4) + float(rng.random())
    return hit, code, depth


@pytest.mark.parametrize("read_lag", [0.0, 0.5, 1.0])
def test_two_ended_scan_equals_brute_force_under_any_interleaving(read_lag):
    rng = np.random.default_rng(1234)
    random.seed(99)
    total_eval = total_unknown = 0
    for _ in range(300):
        n = int(rng.choice([256, 250, 33, 32, 1]))
        hit, code, depth = make_case(rng, n)
        ray = Ray(hit, code, depth)
        nseg = (n + SEG - 1) // SEG
        run_launch([forward_thread(ray, g, read_lag) for g in range(nseg)])
        run_launch([backward_thread(ray, g, read_lag) for g in range(nseg)])
        if hit.any():
            assert ray.bnear == depth[hit].min() and ray.bfar == depth[hit].max()
        else:
            assert ray.bnear == INF and ray.bfar == -1.0    # bound_finish_kernel keeps the sphere near / far then
        total_eval += ray.evaluated
        total_unknown += int((code == 0).sum())
    assert total_eval <= total_unknown      # never more evaluations than the plain scan of the unknown samples


def test_in_order_schedule_skips_the_interior():
    """With the grid order of the launches (segments ascending, then descending) a ray through the object evaluates only
    the unknown samples in front of its first hit and behind its last hit."""
    n = 256
    hit = np.zeros(n, bool)
    hit[60:90] = True        # entry shell
    hit[170:200] = True      # exit shell
    code = np.zeros(n, int)  # no certificate at all: every sample would be evaluated by the plain scan
    depth = np.arange(n, dtype=np.float32) * 0.0078 + 0.5
    ray = Ray(hit, code, depth)
    for g in range(n // SEG):                       # one thread after the other, ascending
        for _ in forward_thread(ray, g, 0.0):
            pass
    for g in reversed(range(n // SEG)):             # descending
        for _ in backward_thread(ray, g, 0.0):
            pass
    assert ray.bnear == depth[60] and ray.bfar == depth[199]
    assert ray.evaluated == 61 + (n - 199)          # samples 0..60 from the front, 255..199 from the back
