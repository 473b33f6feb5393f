"""The schedule of rk_quad_kernel (mossco_code_b200/csrc/msed_rkquad.cuh) as a plain numpy model, runnable without a GPU.

The kernel evaluates the four stages of a Runge-Kutta call (solver_library.F90:142-185) in one walk down the column:
in iteration k stage s works on layer k - (s - 1), reads its own layer from a one-layer link slot (written one
iteration earlier), the layer below from the stage in front of it (written in this iteration), the base state from
the input ring, and the weighted sums of the k_i from the hand-over of the previous iteration.  This test runs that
schedule -- same iteration order, same peeled head and tail, same hand-over order, a ring of six slots, in-place
output three layers behind -- on a nearest-neighbour column RHS and compares it bit for bit with the staged
formulas (whole-array stages, as the Fortran writes them).  It pins the index logic of the walk; the CUDA kernel
itself is compared with the staged path on the GPU (tests/test_gpu_fusion.py, mode quad).
"""
import numpy as np
import pytest

NV = 3


def _rhs_layer(up, cc, dn, coef, k, K):
    """RHS of one layer from its own state, the layer above (flux carried) and below: diffusion + a local reaction."""
    f_up = 0.0 if up is None else coef[k - 1] * (cc - up)          # flux through the upper interface (carried)
    f_dn = 0.0 if dn is None else coef[k] * (dn - cc)              # flux through the lower interface
    react = np.array([-0.3 * cc[0] * cc[1], 0.2 * cc[0] - 0.1 * cc[1], 0.05 * cc[2] * cc[0]])
    return (f_dn - f_up) + react


def _rhs_column(c, coef):
    K = c.shape[0]
    out = np.empty_like(c)
    for k in range(K):
        out[k] = _rhs_layer(c[k - 1] if k > 0 else None, c[k], c[k + 1] if k < K - 1 else None, coef, k, K)
    return out


def _staged(c, coef, dt, is38):
    third = 1.0 / 3.0
    if not is38:                                                   # :147-160
        k1 = _rhs_column(c, coef); c1 = c + 0.5 * dt * k1; acc = 0.5 * k1
        k2 = _rhs_column(c1, coef); c1 = c + 0.5 * dt * k2; acc = acc + k2
        k3 = _rhs_column(c1, coef); c1 = c + dt * k3; acc = acc + k3
        k4 = _rhs_column(c1, coef)
        return c + dt * third * (acc + 0.5 * k4)
    k1 = _rhs_column(c, coef); c1 = c + third * dt * k1            # :169-182
    k2 = _rhs_column(c1, coef); c1 = c + dt * (k2 - third * k1); P = k1 - k2; Q = k1 + 3.0 * k2
    k3 = _rhs_column(c1, coef); c1 = c + dt * (P + k3); Q = Q + 3.0 * k3
    k4 = _rhs_column(c1, coef)
    return c + dt * 1.0 / 8.0 * (Q + k4)


def _walk(c_in, coef, dt, is38):
    """One call as rk_quad_kernel schedules it.  `conc` is updated in place, three layers behind the read front."""
    K = c_in.shape[0]
    assert K >= 5
    conc = c_in.copy()
    third = 1.0 / 3.0
    RING = 6
    ring = [None] * RING                      # input ring: slot = layer mod 6
    link = {2: None, 3: None, 4: None}        # one-layer link slots: the layer stage s evaluates next
    up = {1: None, 2: None, 3: None, 4: None}  # the layer above each stage's current one (stands for the carried flux)
    x12 = x23 = x23b = x34 = None
    fetched = 0

    def fetch():
        nonlocal fetched
        if fetched < K:
            ring[fetched % RING] = conc[fetched].copy()   # reads global memory at issue time
        fetched += 1

    for _ in range(RING - 4):                 # layers 0 and 1
        fetch()
    for k in range(K + 3):
        m = [0, 0, 0, 0]                      # what each stage does: 0 nothing, 1 first layer, 2 inner, 3 last
        for s in range(4):
            layer = k - s
            if 0 <= layer < K:
                m[s] = 1 if layer == 0 else (3 if layer == K - 1 else 2)
        c2, c3, c4 = link[2], link[3], link[4]                    # read before this iteration's results replace them
        rhs1 = rhs2 = rhs3 = None
        y2n = y3n = y4n = None
        if m[0]:
            fetch()                           # layer k+2 into the slot layer k-4 has left
            cc = ring[k % RING]
            cn = ring[(k + 1) % RING] if m[0] != 3 else None
            rhs1 = _rhs_layer(up[1], cc, cn, coef, k, K)
            up[1] = cc
            y2n = cc + (third * dt if is38 else 0.5 * dt) * rhs1
            link[2] = y2n
        if m[1]:
            base = ring[(k - 1) % RING]
            rhs2 = _rhs_layer(up[2], c2, y2n if m[1] != 3 else None, coef, k - 1, K)
            up[2] = c2
            y3n = base + dt * (rhs2 - third * x12) if is38 else base + 0.5 * dt * rhs2
            link[3] = y3n
        if m[2]:
            base = ring[(k - 2) % RING]
            rhs3 = _rhs_layer(up[3], c3, y3n if m[2] != 3 else None, coef, k - 2, K)
            up[3] = c3
            y4n = base + dt * (x23 + rhs3) if is38 else base + dt * rhs3
            link[4] = y4n
        if m[3]:
            base = ring[(k - 3) % RING]
            rhs4 = _rhs_layer(up[4], c4, y4n if m[3] != 3 else None, coef, k - 3, K)
            up[4] = c4
            conc[k - 3] = base + dt * 1.0 / 8.0 * (x34 + rhs4) if is38 else base + dt * third * (x34 + 0.5 * rhs4)
        # hand-over, last stage first: each sum is consumed before the stage in front overwrites it
        if not is38:
            if m[2]: x34 = x23 + rhs3
            if m[1]: x23 = x12 + rhs2
            if m[0]: x12 = 0.5 * rhs1
        else:
            if m[2]: x34 = x23b + 3.0 * rhs3
            if m[1]:
                x23, x23b = x12 - rhs2, x12 + 3.0 * rhs2
            if m[0]: x12 = rhs1
    return conc


@pytest.mark.parametrize("is38", [False, True])
@pytest.mark.parametrize("K", [5, 6, 7, 12, 40])
def test_quad_walk_equals_staged_formulas(K, is38):
    rng = np.random.default_rng(100 + K)
    c = 0.5 + rng.random((K, NV))
    coef = 0.05 + 0.1 * rng.random(K)         # interface coefficient below layer k
    dt = 0.37
    want = _staged(c, coef, dt, is38)
    got = _walk(c, coef, dt, is38)
    # same operations on the same operands: identical up to the association of the scalar factors
    assert np.max(np.abs(got - want) / np.abs(want)) < 5e-15
    # several calls in a row, in place
    a, b = c.copy(), c.copy()
    for _ in range(5):
        a = _staged(a, coef, dt, is38)
        b = _walk(b, coef, dt, is38)
    assert np.max(np.abs(a - b) / np.abs(a)) < 1e-13


def test_quad_walk_reads_every_layer_before_it_is_overwritten():
    """In-place safety: layer k-3 is rewritten in iteration k, the ring fetched it in iteration k-5 at the latest."""
    K = 9
    c = np.arange(K * NV, dtype=float).reshape(K, NV) + 1.0
    coef = np.full(K, 0.1)
    ref = _staged(c, coef, 0.01, False)
    got = _walk(c, coef, 0.01, False)
    assert np.allclose(got, ref, rtol=1e-14, atol=0)
