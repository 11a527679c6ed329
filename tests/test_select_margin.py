"""CPU check of the claim behind select_event_fast (latticemontecarlo_b200/csrc/kmc_kernels.cuh): the one-pass event
selection of the latency kernel -- lane i adds the rates of slots 0..i as a tree, P_i, and compares with u2 * T -- picks
the slot the reference's three sequential passes pick (total, p_q = rate_q / total, running sum; KineticMcFirstOmp.cpp:55-77,
KineticMcAbstract.cpp:106-116) whenever |P_i - u2 T| > 1e-12 T for every slot.  Restated in float64 numpy with the
kernel's summation trees; the GPU test test_one_pass_select_equals_the_sequential_select compares the two kernels' paths."""
import numpy as np

MARGIN = 1e-12


def sequential_select(rates, u2):
    total = np.zeros(len(rates))
    for q in range(12):                                   # left to right, like the reference
        total = total + rates[:, q]
    p = rates / total[:, None]
    c = np.zeros_like(rates)
    run = np.zeros(len(rates))
    for q in range(12):
        run = run + p[:, q]
        c[:, q] = run
    hit = ~(c < u2[:, None])
    return np.where(hit.any(axis=1), hit.argmax(axis=1), 11), c


def one_pass_select(rates, u2):
    P = np.empty_like(rates)
    for i in range(12):                                   # the kernel's tree over the masked rates of lane i
        m = np.where(np.arange(12) <= i, rates, 0.0)
        P[:, i] = ((m[:, 0] + m[:, 1]) + (m[:, 2] + m[:, 3])) + ((m[:, 4] + m[:, 5]) + (m[:, 6] + m[:, 7])) + ((m[:, 8] + m[:, 9]) + (m[:, 10] + m[:, 11]))
    T = P[:, 11]
    d = P - u2[:, None] * T[:, None]
    sure = (np.abs(d) > MARGIN * T[:, None]).all(axis=1)
    hit = ~(d < 0.0)
    return np.where(hit.any(axis=1), hit.argmax(axis=1), 11), sure


def _rates(rng, n):
    # barriers between 0.2 and 1.6 eV at 300..700 K: rates spread over up to 20 decades, some exactly equal
    ea = rng.uniform(0.2, 1.6, (n, 12))
    ea[rng.random((n, 12)) < 0.1] = 0.7
    beta = 1.0 / (8.617333262e-5 * rng.uniform(300.0, 700.0, n))
    return np.exp(-ea * beta[:, None])


def test_one_pass_select_agrees_with_sequential_select_outside_the_margin():
    rng = np.random.default_rng(1)
    n = 400000
    rates = _rates(rng, n)
    u2 = rng.random(n)
    ref, _ = sequential_select(rates, u2)
    fast, sure = one_pass_select(rates, u2)
    assert sure.mean() > 0.999999                          # a uniform u2 practically never lands within 1e-12 of a boundary
    assert np.array_equal(ref[sure], fast[sure])


def test_one_pass_select_near_the_boundaries():
    """u2 placed on, next to and just outside the margin of a cumulative probability: inside the margin the kernel takes
    the sequential path (`sure` is false); from 2e-12 away the two selections must agree."""
    rng = np.random.default_rng(2)
    n = 200000
    rates = _rates(rng, n)
    _, c = sequential_select(rates, np.zeros(n))
    slot = rng.integers(0, 12, n)
    boundary = c[np.arange(n), slot]
    for offset in (0.0, 1e-16, -1e-16, 3e-15, -3e-15, 1e-13, -1e-13):
        u2 = np.clip(boundary + offset, 0.0, np.nextafter(1.0, 0.0))
        ref, _ = sequential_select(rates, u2)
        fast, sure = one_pass_select(rates, u2)
        assert np.array_equal(ref[sure], fast[sure]), offset
        assert sure.mean() < 0.5, offset                   # these u2 sit inside the margin (unless clipped at the ends)
    for offset in (2e-12, -2e-12, 1e-9, -1e-9):
        u2 = np.clip(boundary + offset, 0.0, np.nextafter(1.0, 0.0))
        ref, _ = sequential_select(rates, u2)
        fast, sure = one_pass_select(rates, u2)
        assert np.array_equal(ref[sure], fast[sure]), offset
        assert sure.mean() > 0.8, offset                   # ... and these outside it (bar neighbouring boundaries of tiny rates): the fast path decides, and decides alike
