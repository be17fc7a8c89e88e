"""GPU parity tests of the mtm path — every compute call goes through the C ABI of libb200mtm.so
(host-pointer entry with numpy operands, device-pointer entry with torch CUDA tensors) and is
compared with the CPU oracle on the same inputs.

Bar: BIT-EXACT on integer-valued inputs (every layout, edge size, kernel family and tile config);
on floating-point data the componentwise tolerance north_star states,
    |C - C_exact| <= c * (K+1) * u * (|A||B| + |C0|),   u = 2^-24 (f32) / 2^-53 (f64),
with c = 2 for the FFMA / DFMA / DMMA kernels and c = 4 for 3xTF32 (C_exact evaluated in fp64 /
extended precision on the host).  The measured max ratio is printed.
"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import LAYOUTS, exact_int, int_matrix, order_of, uniform_matrix

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

F32_VARIANTS = ["simt", "3xtf32"]
F64_VARIANTS = ["dfma", "dmma"]
TOL_C = {"simt": 2.0, "dfma": 2.0, "dmma": 2.0, "3xtf32": 4.0}


def variants_for(ob, dtype):
    is64 = np.dtype(dtype) == np.float64
    return [v for v in (F64_VARIANTS if is64 else F32_VARIANTS) if ob.num_configs(v, is64) > 0]


def all_variant_params():
    return [(np.float32, v) for v in F32_VARIANTS] + [(np.float64, v) for v in F64_VARIANTS]


def skip_if_absent(ob, dtype, variant):
    if ob.num_configs(variant, np.dtype(dtype) == np.float64) == 0:
        pytest.skip(f"{variant} not built")


def to_dev(x):
    """numpy 2-D array (any strides) -> torch CUDA tensor with the same logical layout."""
    import torch
    if x.flags["C_CONTIGUOUS"]:
        return torch.from_numpy(x).cuda()
    if x.flags["F_CONTIGUOUS"]:
        return torch.from_numpy(np.ascontiguousarray(x.T)).cuda().t()
    raise ValueError("to_dev expects a contiguous array")


def run_dev(ob, c, a, b, variant, config=None, calls=1):
    import torch
    tc, ta, tb = to_dev(c), to_dev(a), to_dev(b)
    fn = ob.mtm(tc, ta, tb, None, variant=variant, config=config)
    for _ in range(calls):
        fn()
    torch.cuda.synchronize()
    return tc.cpu().numpy()


def tol_bound(c0, a, b, dtype, cfac):
    K = a.shape[1]
    u = np.finfo(dtype).eps / 2
    ab = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    return cfac * (K + 1) * u * (ab + np.abs(c0).astype(np.float64)) + 1e-300


def exact_f(c0, a, b):
    return c0.astype(np.longdouble) + (a.astype(np.longdouble) @ b.astype(np.longdouble))


# ------------------------------------------------------------------------------------------------
# 1. The reference's own test cases (test/test.mtm.cpp): 8 layouts x 2 dtypes x sz 2..31
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,variant", all_variant_params())
@pytest.mark.parametrize("layout", LAYOUTS)
def test_reference_cases_host_entry(layout, dtype, variant, ob, oracle_lib):
    skip_if_absent(ob, dtype, variant)
    rng = np.random.default_rng(abs(hash((layout, variant))) % 2**32)
    for sz in range(2, 32):
        a = int_matrix(rng, (sz, sz), dtype, layout[1])
        b = int_matrix(rng, (sz, sz), dtype, layout[2])
        c = np.zeros((sz, sz), dtype=dtype, order=order_of(layout[0]))
        want = c.copy(order="K")
        oracle_lib.mtm(want, a, b)
        ob.mtm(c, a, b, None, variant=variant)()
        assert np.array_equal(c, want), f"{layout} {variant} sz={sz}"
        assert np.array_equal(c.astype(np.int64), exact_int(np.zeros((sz, sz)), a, b, oracle_lib))


@pytest.mark.parametrize("dtype,variant", all_variant_params())
@pytest.mark.parametrize("layout", LAYOUTS)
def test_reference_cases_device_entry(layout, dtype, variant, ob, oracle_lib):
    """Same cases with device-resident operands: odd leading dimensions reach the kernels
    unpadded (scalar loaders, unaligned epilogue)."""
    skip_if_absent(ob, dtype, variant)
    rng = np.random.default_rng(abs(hash((layout, variant, "dev"))) % 2**32)
    for sz in list(range(2, 32, 3)) + [31]:
        a = int_matrix(rng, (sz, sz), dtype, layout[1])
        b = int_matrix(rng, (sz, sz), dtype, layout[2])
        c0 = int_matrix(rng, (sz, sz), dtype, layout[0])
        want = c0.copy(order="K")
        oracle_lib.mtm(want, a, b)
        got = run_dev(ob, c0, a, b, variant)
        assert np.array_equal(got, want), f"{layout} {variant} sz={sz}"


# ------------------------------------------------------------------------------------------------
# 2. Edge sizes around tile boundaries, rectangular, every tile config and loader combination
# ------------------------------------------------------------------------------------------------
EDGE_SHAPES = [(1, 1, 1), (1, 130, 7), (129, 1, 33), (127, 128, 129), (128, 128, 8), (255, 257, 65),
               (257, 129, 300), (64, 64, 1711), (300, 5, 1030), (3, 515, 260)]


@pytest.mark.parametrize("dtype,variant", all_variant_params())
@pytest.mark.parametrize("layout", ["LLL", "FFF", "FLF", "LFL"])
def test_edge_shapes_bit_exact(layout, dtype, variant, ob, oracle_lib):
    skip_if_absent(ob, dtype, variant)
    rng = np.random.default_rng(abs(hash((layout, variant, "edge"))) % 2**32)
    for (M, N, K) in EDGE_SHAPES:
        a = int_matrix(rng, (M, K), dtype, layout[1])
        b = int_matrix(rng, (K, N), dtype, layout[2])
        c0 = int_matrix(rng, (M, N), dtype, layout[0])
        want = c0.copy(order="K")
        oracle_lib.mtm(want, a, b)
        got_h = c0.copy(order="K")
        ob.mtm(got_h, a, b, None, variant=variant)()
        assert np.array_equal(got_h, want), f"host {layout} {variant} {(M, N, K)}"
        got_d = run_dev(ob, c0, a, b, variant)
        assert np.array_equal(got_d, want), f"dev {layout} {variant} {(M, N, K)}"


@pytest.mark.parametrize("dtype,variant", all_variant_params())
def test_every_tile_config_and_loader(dtype, variant, ob, oracle_lib):
    """Each tile config x {vector-along-mn, vector-along-k, scalar} loaders, via the device entry."""
    skip_if_absent(ob, dtype, variant)
    is64 = np.dtype(dtype) == np.float64
    rng = np.random.default_rng(11)
    seen = set()
    for cfg in range(ob.num_configs(variant, is64)):
        for layout in ("LLL", "LFL", "LLF", "LFF", "FFF"):
            for (M, N, K) in ((264, 392, 136), (261, 387, 131)):   # aligned -> vector loaders; odd -> scalar
                a = int_matrix(rng, (M, K), dtype, layout[1])
                b = int_matrix(rng, (K, N), dtype, layout[2])
                c0 = int_matrix(rng, (M, N), dtype, layout[0])
                want = c0.copy(order="K")
                oracle_lib.mtm(want, a, b)
                got = run_dev(ob, c0, a, b, variant, config=cfg)
                ch = ob.last_choice()
                seen.add((ch["name"], ch["a_mode"], ch["b_mode"]))
                # (a fused 3xTF32 config hands operands it cannot fetch in place to its plane-fed sibling)
                assert ch["config"] == cfg or (variant == "3xtf32" and 2 in (ch["a_mode"], ch["b_mode"])), ch
                assert np.array_equal(got, want), f"{variant} cfg={cfg} {layout} {(M, N, K)} {ch}"
    # every loader combination was exercised; for 3xtf32 the modes are the operand feeds: 0 = in place K-major
    # (TMA straight from the caller's matrix), 1 = in place MN-major, 2 = packed hi/lo planes
    modes = {(a, b) for _, a, b in seen}
    assert {(0, 0), (0, 1), (1, 0), (1, 1), (2, 2)} <= modes, modes


# ------------------------------------------------------------------------------------------------
# 3. Floating-point tolerance
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,variant", all_variant_params())
@pytest.mark.parametrize("layout,shape", [("LLL", (384, 320, 1000)), ("FLF", (257, 130, 4096)),
                                          ("LFL", (96, 200, 2500)), ("FFF", (512, 512, 512))])
def test_tolerance_uniform(layout, shape, dtype, variant, ob, oracle_lib):
    skip_if_absent(ob, dtype, variant)
    M, N, K = shape
    rng = np.random.default_rng(M + 7 * N + 13 * K)
    a = uniform_matrix(rng, (M, K), dtype, layout[1])
    b = uniform_matrix(rng, (K, N), dtype, layout[2])
    c0 = uniform_matrix(rng, (M, N), dtype, layout[0])
    got = c0.copy(order="K")
    ob.mtm(got, a, b, None, variant=variant)()
    exact = exact_f(c0, a, b)
    bound = tol_bound(c0, a, b, dtype, 1.0)
    ratio = float(np.max(np.abs(got.astype(np.longdouble) - exact) / bound))
    print(f"\n[tolerance] {variant} {np.dtype(dtype).name} {layout} {shape}: max |err| / ((K+1) u (|A||B|+|C0|)) = {ratio:.4f}"
          f" (limit {TOL_C[variant]})")
    assert ratio <= TOL_C[variant]
    # and it agrees with the reference algorithm (oracle) to the sum of both error budgets
    ref = c0.copy(order="K")
    oracle_lib.mtm(ref, a, b)
    assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= (TOL_C[variant] + 2.0) * bound)


@pytest.mark.parametrize("dtype,variant", all_variant_params())
def test_tolerance_wide_dynamic_range(dtype, variant, ob):
    """Exponentially distributed magnitudes with mixed signs (SURVEY 8d input set iii)."""
    skip_if_absent(ob, dtype, variant)
    rng = np.random.default_rng(99)
    M, N, K = 200, 264, 1536
    span = 12 if np.dtype(dtype) == np.float32 else 40
    mk = lambda shape: (rng.choice([-1.0, 1.0], shape) * np.exp2(rng.uniform(-span, span, shape))).astype(dtype)
    a, b, c0 = mk((M, K)), np.asfortranarray(mk((K, N))), mk((M, N))
    got = c0.copy()
    ob.mtm(got, a, b, None, variant=variant)()
    exact = exact_f(c0, a, b)
    ratio = float(np.max(np.abs(got.astype(np.longdouble) - exact) / tol_bound(c0, a, b, dtype, 1.0)))
    print(f"\n[tolerance/wide] {variant} {np.dtype(dtype).name}: ratio {ratio:.4f} (limit {TOL_C[variant]})")
    assert ratio <= TOL_C[variant]


def test_golden_vectors_gpu(ob):
    """Outputs of the unmodified reference (tests/golden/mtm_golden.npz) within the stated tolerance
    (rounding order differs from the CPU blocking, so not bit-wise)."""
    g = np.load(Path(__file__).parent / "golden" / "mtm_golden.npz")
    for i in range(int(g["ncases"])):
        a, b, c0, want = g[f"case{i}_a"], g[f"case{i}_b"], g[f"case{i}_c0"], g[f"case{i}_c"]
        for variant in variants_for(ob, a.dtype):
            got = c0.copy(order="K")
            ob.mtm(got, a, b, None, variant=variant)()
            bound = tol_bound(c0, a, b, a.dtype, TOL_C[variant] + 2.0)
            assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= bound), (i, variant)


def _dev_random(shape, dtype, layout, seed, kind):
    """uniform(-1, 1) or wide-dynamic-range (sign * 2^uniform(-span, span)) matrix generated on the device."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    shp = shape if layout == "L" else (shape[1], shape[0])
    if kind == "uniform":
        x = torch.rand(shp, device="cuda", dtype=tdt, generator=g) * 2 - 1
    else:
        span = 12.0 if tdt == torch.float32 else 40.0
        mag = torch.exp2((torch.rand(shp, device="cuda", dtype=tdt, generator=g) * 2 - 1) * span)
        x = mag * (torch.randint(0, 2, shp, device="cuda", generator=g).to(tdt) * 2 - 1)
    return x if layout == "L" else x.t()


BIG_TOL = [  # BASELINE.json sizes: (name, dtype, variants, (M, N, K), layout)
    ("config2 8192^3 f32 LLL", np.float32, F32_VARIANTS, (8192, 8192, 8192), "LLL"),
    ("config3 8192^3 f64 LLL", np.float64, F64_VARIANTS, (8192, 8192, 8192), "LLL"),
    ("config2 16384^3 f32 LLL", np.float32, F32_VARIANTS, (16384, 16384, 16384), "LLL"),
    ("config4 65536x1024x1024 f32 LLL", np.float32, F32_VARIANTS, (65536, 1024, 1024), "LLL"),
    ("config4 65536x1024x1024 f32 FLF", np.float32, F32_VARIANTS, (65536, 1024, 1024), "FLF"),
]


@pytest.mark.parametrize("kind", ["uniform", "wide"])
@pytest.mark.parametrize("name,dtype,variants,shape,layout", BIG_TOL, ids=[f[0] for f in BIG_TOL])
def test_tolerance_at_baseline_sizes(name, dtype, variants, shape, layout, kind, ob):
    """north_star's componentwise tolerance at the sizes BASELINE.json quotes, on non-integer data (the
    integer full-size tests cannot see 3xTF32's lo planes or fp32 accumulation error): 96 sampled rows x all
    columns against an fp64 product on the device; the measured ratio is printed."""
    import torch
    M, N, K = shape
    a = _dev_random((M, K), dtype, layout[1], 11, kind)
    b = _dev_random((K, N), dtype, layout[2], 12, kind)
    c0 = _dev_random((M, N), dtype, layout[0], 13, kind)
    rows = torch.unique(torch.linspace(0, M - 1, 96, device="cuda").long())
    bd = b.double()
    exact = c0[rows].double() + a[rows].double() @ bd
    u = float(np.finfo(dtype).eps) / 2
    bound = (K + 1) * u * (a[rows].double().abs() @ bd.abs() + c0[rows].double().abs()) + 1e-300
    del bd
    for variant in variants:
        if ob.num_configs(variant, np.dtype(dtype) == np.float64) == 0:
            continue
        c = c0.clone(memory_format=torch.preserve_format)
        ob.mtm(c, a, b, None, variant=variant)()
        torch.cuda.synchronize()
        got = c[rows].double()
        assert bool(torch.isfinite(got).all()), f"{name} {variant} {kind}: non-finite result"
        ratio = float(((got - exact).abs() / bound).max().item())
        print(f"\n[tolerance@baseline] {name} {variant} {kind}: max |err| / ((K+1) u (|A||B|+|C0|)) = {ratio:.4f} "
              f"(limit {TOL_C[variant]}), kernel {ob.last_choice()['name']}")
        assert ratio <= TOL_C[variant], f"{name} {variant} {kind}"
        del c
    del a, b, c0
    torch.cuda.empty_cache()


@pytest.mark.parametrize("variant", F32_VARIANTS)
def test_extreme_magnitudes_fp32(variant, ob):
    """Operand split at the ends of the fp32 range.  |x| within 2^-11 of FLT_MAX: a round-to-nearest TF32 hi
    would overflow to Inf; the truncating split must stay finite and within tolerance.
    Subnormal inputs: reported (the tensor pipe may flush them; the FFMA path may not)."""
    import torch
    skip_if_absent(ob, np.float32, variant)
    M = N = K = 256
    big = float(np.float32(np.finfo(np.float32).max))
    x = torch.full((M, K), big, device="cuda")
    x[::2] = torch.nextafter(torch.tensor(big), torch.tensor(0.0)).item()
    x[:, 1::2] *= -1
    b = torch.zeros((K, N), device="cuda")
    b.fill_diagonal_(2.0 ** -10)                       # exactly one non-zero product per output: no overflow in the sum
    c = torch.zeros((M, N), device="cuda")
    ob.mtm(c, x, b, None, variant=variant, config=(1 if variant == "3xtf32" else None))()
    torch.cuda.synchronize()
    want = x.double() * 2.0 ** -10
    assert bool(torch.isfinite(c).all()), f"{variant}: non-finite result near FLT_MAX"
    rel = float(((c.double() - want).abs() / want.abs()).max().item())
    print(f"\n[extreme] {variant} near FLT_MAX: max relative error {rel:.3e}")
    assert rel <= 4 * (K + 1) * 2.0 ** -24
    # a non-finite operand makes exactly its row non-finite (FFMA: Inf; 3xTF32: Inf * lo(b) with lo(b) == 0 is NaN,
    # as in every split scheme) and leaves the other rows exact
    x2 = torch.ones((M, K), device="cuda")
    x2[3, 5] = float("inf")
    b2 = torch.ones((K, N), device="cuda")
    c2 = torch.zeros((M, N), device="cuda")
    ob.mtm(c2, x2, b2, None, variant=variant, config=(1 if variant == "3xtf32" else None))()
    torch.cuda.synchronize()
    assert not bool(torch.isfinite(c2[3]).any()), f"{variant}: Inf row became {c2[3, :4].tolist()}"
    if variant != "3xtf32":
        assert bool(torch.isinf(c2[3]).all()) and bool((c2[3] > 0).all())
    assert bool((c2[:3] == K).all()) and bool((c2[4:] == K).all())
    # subnormals: report what the path does with them
    sub = torch.full((M, K), 2.0 ** -140, device="cuda")
    b3 = torch.zeros((K, N), device="cuda")
    b3.fill_diagonal_(2.0 ** 100)
    c3 = torch.zeros((M, N), device="cuda")
    ob.mtm(c3, sub, b3, None, variant=variant, config=(1 if variant == "3xtf32" else None))()
    torch.cuda.synchronize()
    v = float(c3[0, 0].item())
    print(f"[extreme] {variant} subnormal input 2^-140 * 2^100 -> {v:.6e} (exact 2^-40 = {2.0 ** -40:.6e}; 0 = flushed)")
    assert v == 0.0 or abs(v - 2.0 ** -40) <= 2.0 ** -40 * 2.0 ** -9


def test_host_slab_pipeline_auto_resolves_once(ob):
    """with AUTO flags every slab of the host pipeline has to run the kernel family and tile
    config the FIRST slab resolved to — a shorter tail slab re-resolving AUTO would switch family (and read a
    B image that family never wrote).  2400 x 16384 x 1024 (C, A row-major, B column-major): slabs of 512 rows
    and a 352-row tail.  The result must equal one unsliced device call of that family bit for bit."""
    import torch
    M, N, K = 2400, 16384, 1024
    rng = np.random.default_rng(17)
    a = uniform_matrix(rng, (M, K), np.float32, "L")
    b = uniform_matrix(rng, (K, N), np.float32, "F")
    c0 = uniform_matrix(rng, (M, N), np.float32, "L")
    got = c0.copy(order="K")
    ob.mtm(got, a, b, None)()                       # AUTO
    ch = ob.last_choice()
    assert ch["launches"] > 2, ch                   # it was sliced
    want = run_dev(ob, c0, a, b, ch["variant"])
    assert np.array_equal(got, want), ch
    exact = c0.astype(np.float64) + a.astype(np.float64) @ b.astype(np.float64)
    ratio = float(np.max(np.abs(got.astype(np.float64) - exact) / tol_bound(c0, a, b, np.float32, 1.0)))
    assert ratio <= TOL_C[ch["variant"]]
    # column-major C: the sliced operand is B, the shared one A
    a2 = uniform_matrix(rng, (N, K), np.float32, "L")
    b2 = uniform_matrix(rng, (K, M), np.float32, "F")
    c2 = np.asfortranarray(uniform_matrix(rng, (N, M), np.float32, "L"))
    got2 = c2.copy(order="K")
    ob.mtm(got2, a2, b2, None)()
    ch2 = ob.last_choice()
    want2 = run_dev(ob, c2, a2, b2, ch2["variant"])
    assert np.array_equal(got2, want2), ch2


# ------------------------------------------------------------------------------------------------
# 4. Semantics: accumulate, sub-views, untouched memory, error codes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,variant", all_variant_params())
def test_accumulates_on_every_call(dtype, variant, ob):
    """src/mtm.cpp:204-208 protocol: all-ones inputs, C never re-zeroed; after n calls C == n*K."""
    skip_if_absent(ob, dtype, variant)
    M = N = K = 320
    a = np.ones((M, K), dtype, order="F")
    b = np.ones((K, N), dtype, order="F")
    c = np.zeros((M, N), dtype, order="F")
    fn = ob.mtm(c, a, b, None, variant=variant)
    for n in range(1, 6):
        fn()
        assert np.all(c == n * K)
    got = run_dev(ob, np.zeros((M, N), dtype), a, b, variant, calls=5)
    assert np.all(got == 5 * K)


def test_dependent_launch_chain_of_mixed_calls(ob):
    """The 3xTF32 split pass and MMA kernel are programmatic dependent launches: the split of call i + 1 is scheduled
    while the MMA kernel of call i still runs and overwrites the planes that kernel reads, so it has to wait for it in the
    kernel.  A chain of calls of DIFFERENT shapes and tile configs (their planes overlap in the shared workspace), issued
    back to back on one stream without any host synchronisation, on integer data: every C must come out exact."""
    import torch
    if ob.num_configs("3xtf32", False) == 0:
        pytest.skip("3xTF32 path not built")
    g = torch.Generator(device="cuda").manual_seed(11)
    cases = []
    for (M, N, K, cfg) in ((512, 1024, 256, 9), (300, 260, 520, None), (1024, 768, 640, 0), (256, 512, 2048, 4), (128, 128, 128, 5),
                           (640, 384, 96, 1), (2048, 1024, 512, 9)):
        if cfg is not None and cfg >= ob.num_configs("3xtf32", False):
            cfg = None
        a = torch.randint(0, 8, (M, K), device="cuda", generator=g).float()
        b = torch.randint(0, 8, (K, N), device="cuda", generator=g).float()
        cases.append((a, b, torch.zeros((M, N), device="cuda"), cfg))
    rounds = 12
    for _ in range(rounds):
        for (a, b, c, cfg) in cases:
            ob.mtm(c, a, b, None, variant="3xtf32", config=cfg)()
    torch.cuda.synchronize()
    for (a, b, c, cfg) in cases:
        want = rounds * (a.double() @ b.double())
        assert torch.equal(c.double(), want), (tuple(a.shape), tuple(b.shape), cfg)


@pytest.mark.parametrize("dtype,variant", all_variant_params())
def test_strided_subviews(dtype, variant, ob, oracle_lib):
    """A and B with both strides != 1 (utils.hpp:99-141 honours both); C a sub-view with ldc > N.
    Elements of the parent buffers outside the views must stay untouched."""
    import torch
    skip_if_absent(ob, dtype, variant)
    rng = np.random.default_rng(5)
    big_a = rng.integers(0, 100, (300, 400)).astype(dtype)
    big_b = rng.integers(0, 100, (500, 330)).astype(dtype)
    big_c = rng.integers(0, 100, (260, 300)).astype(dtype)
    sa = np.s_[3:243:2, 1:391:3]     # 120 x 130, strides (800, 3)
    sb = np.s_[5:395:3, 2:302:2]     # 130 x 150, strides (990, 2)
    sc = np.s_[7:127, 11:161]        # 120 x 150, row-major sub-view
    want = big_c.copy()
    oracle_lib.mtm(want[sc], big_a[sa], big_b[sb])
    # host entry
    got = big_c.copy()
    ob.mtm(got[sc], big_a[sa], big_b[sb], None, variant=variant)()
    assert np.array_equal(got, want)
    # device entry (torch views carry the strides)
    ta, tb, tc = torch.from_numpy(big_a).cuda(), torch.from_numpy(big_b).cuda(), torch.from_numpy(big_c).cuda()
    ob.mtm(tc[7:127, 11:161], ta[3:243:2, 1:391:3], tb[5:395:3, 2:302:2], None, variant=variant)()
    torch.cuda.synchronize()
    assert np.array_equal(tc.cpu().numpy(), want)
    # column-major C sub-view
    big_cf = np.asfortranarray(big_c)
    want_f = big_cf.copy(order="K")
    oracle_lib.mtm(want_f[sc], big_a[sa], big_b[sb])
    got_f = big_cf.copy(order="K")
    ob.mtm(got_f[sc], big_a[sa], big_b[sb], None, variant=variant)()
    assert np.array_equal(got_f, want_f)


@pytest.mark.parametrize("dtype,variant", all_variant_params())
@pytest.mark.parametrize("layout", LAYOUTS)
def test_host_slab_pipeline_matches_device_entry(layout, dtype, variant, ob):
    """Host-pointer calls above 48 MB are cut into slabs along C's slow dimension and pipelined over
    three streams (mtm_api.cu: mtm_host); each slab is an ordinary device call, so the result must be
    bit-identical to one unsliced device call — also on non-integer data — for row- and column-
    contiguous C, and the operand shared by all slabs must be re-laid only once (reuse_b)."""
    import torch
    skip_if_absent(ob, dtype, variant)
    M, N, K = (3000, 2900, 1400) if np.dtype(dtype) == np.float32 else (2100, 2000, 1100)
    rng = np.random.default_rng(3)
    a = uniform_matrix(rng, (M, K), dtype, layout[1])
    b = uniform_matrix(rng, (K, N), dtype, layout[2])
    c0 = uniform_matrix(rng, (M, N), dtype, layout[0])
    assert a.nbytes + b.nbytes + c0.nbytes >= 48 * 1024 * 1024
    got = c0.copy(order="K")
    ob.mtm(got, a, b, None, variant=variant)()
    launches_host = ob.last_choice()["launches"]
    want = run_dev(ob, c0, a, b, variant)
    launches_dev = ob.last_choice()["launches"]
    assert np.array_equal(got, want), f"{layout} {variant}"
    assert launches_host >= launches_dev            # several slabs were launched
    if np.dtype(dtype) == np.float32:               # (the device entry's own accuracy is pinned by test_tolerance_*)
        exact = c0.astype(np.float64) + a.astype(np.float64) @ b.astype(np.float64)
        ratio = float(np.max(np.abs(got.astype(np.float64) - exact) / tol_bound(c0, a, b, dtype, 1.0)))
        assert ratio <= TOL_C[variant]


@pytest.mark.parametrize("dtype,variant", all_variant_params())
def test_randomized_differential(dtype, variant, ob, oracle_lib):
    """Seeded random shapes (1..300), layouts, sub-view strides and offsets, host and device entry,
    integer-valued data: bit-exact against the oracle, and nothing outside the C view is touched."""
    import torch
    skip_if_absent(ob, dtype, variant)
    rng = np.random.default_rng(20261017)
    tdev = lambda x: torch.from_numpy(x).cuda()
    for case in range(40):
        M, N, K = (int(v) for v in rng.integers(1, 301, 3))
        def view(rows, cols):
            """A (rows x cols) view of a larger row- or column-major parent with random steps/offsets."""
            sr, sc = int(rng.integers(1, 3)), int(rng.integers(1, 3))
            o0, o1 = int(rng.integers(0, 4)), int(rng.integers(0, 4))
            parent = rng.integers(0, 100, (o0 + rows * sr + 2, o1 + cols * sc + 3)).astype(dtype)
            if rng.integers(0, 2):
                parent = np.asfortranarray(parent)
            sl = np.s_[o0:o0 + rows * sr:sr, o1:o1 + cols * sc:sc]
            return parent, sl
        pa, sa = view(M, K)
        pb, sb = view(K, N)
        # C: unit stride in one dimension (the reference's requirement), arbitrary leading dimension/offset
        o0, o1 = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        pc = rng.integers(0, 100, (o0 + M + 2, o1 + N + 3)).astype(dtype)
        if rng.integers(0, 2):
            pc = np.asfortranarray(pc)
        sc_ = np.s_[o0:o0 + M, o1:o1 + N]
        want = pc.copy(order="K")
        oracle_lib.mtm(want[sc_], pa[sa], pb[sb])
        got = pc.copy(order="K")
        ob.mtm(got[sc_], pa[sa], pb[sb], None, variant=variant)()
        assert np.array_equal(got, want), ("host", case, M, N, K)
        def dev_parent(x):      # keep the parent's memory order on the device
            return tdev(x) if x.flags["C_CONTIGUOUS"] else tdev(np.ascontiguousarray(x.T)).t()
        ta, tb, tc = dev_parent(pa), dev_parent(pb), dev_parent(pc)
        ob.mtm(tc[sc_], ta[sa], tb[sb], None, variant=variant)()
        torch.cuda.synchronize()
        assert np.array_equal(tc.cpu().numpy(), want), ("dev", case, M, N, K, ob.last_choice())


def test_c_abi_status_codes_on_gpu(ob):
    import ctypes
    L = ob.lib()
    S2 = ctypes.c_size_t * 2
    buf = (ctypes.c_float * 64)()
    assert L.b200_mtm_f32(buf, S2(4, 3), S2(3, 1), buf, S2(4, 5), S2(5, 1), buf, S2(6, 3), S2(3, 1), 0) == 2
    assert L.b200_mtm_f32(buf, S2(4, 3), S2(6, 2), buf, S2(4, 5), S2(5, 1), buf, S2(5, 3), S2(3, 1), 0) == 3
    assert L.b200_mtm_f32(buf, S2(4, 3), S2(3, 1), buf, S2(4, 5), S2(5, 1), buf, S2(5, 3), S2(3, 1), 3) == 1  # DFMA on f32
    assert L.b200_mtm_f32(buf, S2(0, 3), S2(3, 1), buf, S2(0, 5), S2(5, 1), buf, S2(5, 3), S2(3, 1), 0) == 0  # empty: no-op
    c = np.full((4, 3), 7.0, np.float32)
    ob.mtm  # K == 0 leaves C unchanged
    rc = L.b200_mtm_f32(c.ctypes.data_as(ctypes.c_void_p), S2(4, 3), S2(3, 1), buf, S2(4, 0), S2(1, 1), buf,
                        S2(0, 3), S2(3, 1), 0)
    assert rc == 0 and np.all(c == 7.0)


def test_launches_are_counted(ob):
    before = ob.launch_count()
    a = np.ones((64, 64), np.float32)
    c = np.zeros((64, 64), np.float32)
    ob.mtm(c, a, a, None, variant="simt")()
    assert ob.launch_count() == before + 1
    assert ob.last_choice()["name"].startswith("ffma_")


# ------------------------------------------------------------------------------------------------
# 5. BASELINE.json's full sizes, through size-independent exact properties
# ------------------------------------------------------------------------------------------------
def _small_int_dev(shape, hi, dtype, layout, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    if layout == "L":
        return torch.randint(0, hi, shape, device="cuda", generator=g).to(tdt)
    return torch.randint(0, hi, (shape[1], shape[0]), device="cuda", generator=g).to(tdt).t()


FULL = [  # (name, dtype, variants, (M,N,K), layout, value range so that K*(hi-1)^2 < 2^24)
    ("config2 8192^3 f32 LLL", np.float32, F32_VARIANTS, (8192, 8192, 8192), "LLL", 10),
    ("config2 4096^3 f32 FFF", np.float32, F32_VARIANTS, (4096, 4096, 4096), "FFF", 50),
    ("config3 8192^3 f64 LLL", np.float64, F64_VARIANTS, (8192, 8192, 8192), "LLL", 100),
    ("config4 65536x1024x1024 f32 LLL", np.float32, F32_VARIANTS, (65536, 1024, 1024), "LLL", 100),
    ("config4 65536x1024x1024 f32 FLF", np.float32, F32_VARIANTS, (65536, 1024, 1024), "FLF", 100),
    ("config2 16384^3 f32 LLL", np.float32, F32_VARIANTS, (16384, 16384, 16384), "LLL", 6),
    ("config5 32768^3 f32 LLL (whole problem on one GPU)", np.float32, F32_VARIANTS, (32768, 32768, 32768), "LLL", 4),
]


@pytest.mark.parametrize("name,dtype,variants,shape,layout,hi", FULL, ids=[f[0] for f in FULL])
def test_full_size_exact(name, dtype, variants, shape, layout, hi, ob):
    """Integer-valued inputs small enough that every partial sum is exact in the compute type:
    any summation order (FFMA, 3xTF32, DFMA, DMMA) must reproduce the exact product, checked on
    the device against fp64 (exact for these magnitudes).  C starts non-zero and is called twice."""
    import torch
    M, N, K = shape
    a = _small_int_dev((M, K), hi, dtype, layout[1], 1)
    b = _small_int_dev((K, N), hi, dtype, layout[2], 2)
    c0 = _small_int_dev((M, N), hi, dtype, layout[0], 3)
    rows = torch.arange(0, M, max(1, M // 512), device="cuda")[:512]        # sampled rows, all columns
    want_rows = c0[rows].double() + 2 * (a[rows].double() @ b.double())
    for variant in variants:
        if ob.num_configs(variant, np.dtype(dtype) == np.float64) == 0:
            continue
        c = c0.clone(memory_format=torch.preserve_format)
        fn = ob.mtm(c, a, b, None, variant=variant)
        fn()
        fn()
        torch.cuda.synchronize()
        assert torch.equal(c[rows].double(), want_rows), f"{name} {variant}"
        # checksum of checksums over the whole matrix: sum(C) == sum(C0) + 2 * colsum(A) . rowsum(B)
        total = c.double().sum()
        want_total = c0.double().sum() + 2 * torch.dot(a.double().sum(0), b.double().sum(1))
        assert total.item() == want_total.item(), f"{name} {variant} checksum"
        del c
    del a, b, c0
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------
# 6. The C++ front-end (include/mtm.hpp), re-expressed test/test.mtm.cpp
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cpp_test_exe(ob, tmp_path_factory):
    exe = tmp_path_factory.mktemp("cpp") / "test_mtm"
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", f"-I{ROOT / 'include' / 'compat'}", f"-I{ROOT / 'include'}",
           str(ROOT / "tests" / "cpp" / "test_mtm.cpp"), "-o", str(exe),
           f"-L{ob.library_path().parent}", "-lb200mtm", f"-Wl,-rpath,{ob.library_path().parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("variant", [0, 1, 2, 4])
def test_cpp_front_end(variant, cpp_test_exe, ob):
    if variant == 2 and ob.num_configs("3xtf32", False) == 0:
        pytest.skip("3xtf32 not built")
    r = subprocess.run([str(cpp_test_exe), str(variant)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_cpp_harness_sweep(ob, tmp_path):
    """tools/mtm_harness.cpp — the src/mtm.cpp protocol (all-ones inputs, accumulating C, flops =
    M*N*(2K-1), metric.str() + csv) on a short sweep, device-resident and host-tensor series."""
    exe = tmp_path / "mtm_harness"
    lib = ob.library_path().parent
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", f"-I{ROOT / 'include' / 'compat'}", f"-I{ROOT / 'include'}",
           str(ROOT / "tools" / "mtm_harness.cpp"), "-o", str(exe), f"-L{lib}", "-lb200mtm", f"-Wl,-rpath,{lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    csv = tmp_path / "tensor.csv"
    for typ in ("f32", "f64"):
        r = subprocess.run([str(exe), "--type", typ, "--layout", "L", "--max", "512", "--step", "96", "--host",
                            "--csv", str(csv)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "Peak Performance:" in r.stdout and "Max Peak Utilization in %" in r.stdout
        lines = csv.read_text().strip().splitlines()
        assert lines[0].count('"') == 6 and "host-tensors" in lines[0]       # three quoted series
        assert len(lines) == 1 + 6                                             # 32, 128, ..., 512
        assert all(float(v) > 0 for l in lines[1:] for v in l.split(","))


# ---- K-panel calls as the multi-GPU drivers issue them ------------------------------------------------
@pytest.mark.parametrize("variant", F32_VARIANTS)
@pytest.mark.parametrize("M,N", [(2048, 4224), (1100, 2176)])
def test_k_panel_sequence_of_strided_views(variant, M, N, ob):
    """C += A[:, k0:k1] * B[k0:k1, :] for consecutive K-panels of different widths, back to back on one
    stream without synchronising in between (sharded.py: RowBlockMtm chunks, SummaMtm panels): A panels
    are column sub-ranges of a wider row-major matrix, and the tensor-core path's workspace layout
    changes from call to call.  Integer-valued data: the sum must be exact."""
    import torch
    skip_if_absent(ob, np.float32, variant)
    K = 4000
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randint(0, 10, (M, K), device="cuda", generator=g).float()
    B = torch.randint(0, 10, (K, N), device="cuda", generator=g).float()
    C0 = torch.randint(0, 10, (M, N), device="cuda", generator=g).float()
    c = C0.clone()
    stage = [torch.empty((1024, N), device="cuda") for _ in range(2)]
    panels = [(0, 1024), (1024, 2016), (2016, 3040), (3040, 4000)]
    for rep in range(2):
        for t, (k0, k1) in enumerate(panels):
            bp = stage[t & 1][: k1 - k0]
            bp.copy_(B[k0:k1])                                   # as a received panel buffer
            ob.mtm(c, A[:, k0:k1], bp, None, variant=variant)()
    torch.cuda.synchronize()
    want = C0.double() + 2 * (A.double() @ B.double())
    bad = (c.double() != want)
    assert not bool(bad.any()), (int(bad.sum()), bad.nonzero()[0].tolist(), bad.nonzero()[-1].tolist())


# ---- the multi-GPU host-pointer entry (b200_mtm_*_mgpu) -------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("layout", LAYOUTS)      # all 8: a column-major A cut by rows / a row-major B cut by columns included
def test_mgpu_entry_matches_single_gpu(layout, dtype, ob, oracle_lib):
    """One call spread over every visible GPU (one on the driver's test box, where the entry must fall through to
    the single-GPU path; 2..8 under `gpurun --gpus N`): bit-exact on integer data against the oracle, and on
    uniform data bit-identical to the single-GPU host call (K is never split: same kernel, same order)."""
    n_dev = ob.device_count()
    M, N, K = (2600, 2300, 1500) if np.dtype(dtype) == np.float32 else (1900, 1700, 1100)
    rng = np.random.default_rng(23)
    a = int_matrix(rng, (M, K), dtype, layout[1])
    b = int_matrix(rng, (K, N), dtype, layout[2])
    c0 = int_matrix(rng, (M, N), dtype, layout[0])
    want = c0.copy(order="K")
    oracle_lib.mtm(want, a, b)
    got = c0.copy(order="K")
    ob.mtm(got, a, b, None, devices=0)()
    assert np.array_equal(got, want), f"{layout} on {n_dev} device(s): {ob.last_choice()}"
    au = uniform_matrix(rng, (M, K), dtype, layout[1])
    bu = uniform_matrix(rng, (K, N), dtype, layout[2])
    cu = uniform_matrix(rng, (M, N), dtype, layout[0])
    one = cu.copy(order="K")
    ob.mtm(one, au, bu, None)()
    ch = ob.last_choice()
    many = cu.copy(order="K")
    ob.mtm(many, au, bu, None, devices=0, variant=ch["variant"], config=ch["config"])()
    assert np.array_equal(one, many), f"{layout} {n_dev} device(s)"
    if n_dev >= 2:      # an explicit device list, in reverse order
        rev = cu.copy(order="K")
        ob.mtm(rev, au, bu, None, devices=list(range(n_dev))[::-1], variant=ch["variant"], config=ch["config"])()
        assert np.array_equal(one, rev)
    print(f"\n[mgpu] {layout} {np.dtype(dtype).name}: {n_dev} device(s), kernel {ch['name']}")


def test_mgpu_entry_rejects_bad_device_lists(ob):
    a = np.ones((600, 64), np.float32)
    b = np.ones((64, 600), np.float32)
    c = np.zeros((600, 600), np.float32)
    with pytest.raises(ob.B200Error):
        ob.mtm(c, a, b, None, devices=[0, 0])()
    with pytest.raises(ob.B200Error):
        ob.mtm(c, a, b, None, devices=[ob.device_count() + 3])()
    ob.mtm(c, a, b, None, devices=[0])()
    assert np.all(c == 64)


# ---- split-K of the tensor-core path (small problems) ---------------------------------------------------
@pytest.mark.parametrize("split_k", [0, 1, 2, 3, 8])
def test_split_k_exact_and_deterministic(split_k, ob, oracle_lib):
    """Few output tiles: K is split over work units whose partial products are added into C in a FIXED order
    (turnstile per tile row-block).  Integer data: exact for every split count and tile config; uniform data:
    repeated calls give the same bits (no dependence on which split finishes first) within the tolerance."""
    import torch
    skip_if_absent(ob, np.float32, "3xtf32")
    rng = np.random.default_rng(31 + split_k)
    for (M, N, K) in ((256, 256, 2048), (300, 260, 1111), (512, 512, 512), (1024, 768, 1024), (130, 2000, 4096)):
        for cfg in (0, 1, 2, 4, 5, 6, 7, 8):
            a = int_matrix(rng, (M, K), np.float32, "L")
            b = int_matrix(rng, (K, N), np.float32, "L")
            b = (b % 10).astype(np.float32)                      # keep K * 99 * 9 below 2^24 for K = 4096
            c0 = int_matrix(rng, (M, N), np.float32, "L")
            want = c0.copy()
            oracle_lib.mtm(want, a, b)
            tc, ta, tb = to_dev(c0), to_dev(a), to_dev(b)
            ob.mtm(tc, ta, tb, None, variant="3xtf32", config=cfg, split_k=split_k)()
            torch.cuda.synchronize()
            assert np.array_equal(tc.cpu().numpy(), want), f"split_k={split_k} cfg={cfg} {(M, N, K)} {ob.last_choice()}"
    M, N, K = 512, 384, 4096
    a = uniform_matrix(rng, (M, K), np.float32, "L")
    b = uniform_matrix(rng, (K, N), np.float32, "F")
    c0 = uniform_matrix(rng, (M, N), np.float32, "L")
    outs = []
    for _ in range(4):
        tc, ta, tb = to_dev(c0), to_dev(a), to_dev(b)
        ob.mtm(tc, ta, tb, None, variant="3xtf32", split_k=split_k)()
        torch.cuda.synchronize()
        outs.append(tc.cpu().numpy())
    name = ob.last_choice()["name"]
    assert all(np.array_equal(outs[0], o) for o in outs[1:]), f"split_k={split_k}: run-to-run differences ({name})"
    exact = exact_f(c0, a, b)
    ratio = float(np.max(np.abs(outs[0].astype(np.longdouble) - exact) / tol_bound(c0, a, b, np.float32, 1.0)))
    print(f"\n[split-K] request {split_k} -> {name}: ratio {ratio:.4f}")
    assert ratio <= TOL_C["3xtf32"]
    if split_k >= 2:
        assert f"splitk" in name


def test_cpp_front_end_spreads_over_all_gpus(ob, tmp_path):
    """tools/mtm_mgpu_check.cpp: ONE amt::mtm call on make_tensor (pageable) host storage, routed by include/mtm.hpp to
    b200_mtm_f32_mgpu when several GPUs are visible (>= 2 * 4096^3 flop), to b200_mtm_f32 otherwise; integer data,
    sampled rows compared with an integer product on the host.  Row-major and column-major tensors."""
    import json
    exe = tmp_path / "mtm_mgpu_check"
    lib = ob.library_path().parent
    cmd = ["/usr/bin/g++", "-std=c++20", "-O2", "-fopenmp", f"-I{ROOT / 'include' / 'compat'}", f"-I{ROOT / 'include'}",
           str(ROOT / "tools" / "mtm_mgpu_check.cpp"), "-o", str(exe), f"-L{lib}", "-lb200mtm", f"-Wl,-rpath,{lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for layout in ("L", "F"):
        r = subprocess.run([str(exe), "--size", "4096", "--calls", "2", "--layout", layout], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        print(f"\n[cpp mgpu] {res}")
        assert res["exact"] and res["mismatches"] == 0
        if res["devices_visible"] > 1:
            assert res["launches"] >= 2 * res["devices_visible"]      # every device ran its shard


_STREAM_K_CHILD = r"""
import sys, json
sys.path.insert(0, %r)
import torch
import openmp_blas_b200 as ob
res = []
for shape, cfg in [((2048, 2048, 2048), 0), ((2048, 2048, 2048), 1), ((4096, 4096, 1024), 0), ((1024, 1024, 1024), 5),
                   ((4096, 2304, 520), 4), ((2048, 2048, 2048), 6)]:
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randint(0, 10, (M, K), device="cuda", generator=g).float()
    b = torch.randint(0, 10, (K, N), device="cuda", generator=g).float()
    c0 = torch.randint(0, 10, (M, N), device="cuda", generator=g).float()
    c = c0.clone()
    fn = ob.mtm(c, a, b, None, variant="3xtf32", config=cfg)
    fn(); fn()
    torch.cuda.synchronize()
    name = ob.last_choice()["name"]
    exact = bool(torch.equal(c.double(), c0.double() + 2 * (a.double() @ b.double())))
    au = torch.rand((M, K), device="cuda", generator=g) * 2 - 1
    bu = (torch.rand((N, K), device="cuda", generator=g) * 2 - 1).t()
    outs = []
    for _ in range(3):
        cu = torch.zeros((M, N), device="cuda")
        ob.mtm(cu, au, bu, None, variant="3xtf32", config=cfg)()
        torch.cuda.synchronize()
        outs.append(cu)
    same = bool(torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]))
    rows = torch.arange(0, M, max(1, M // 64), device="cuda")
    ex = au[rows].double() @ bu.double()
    bound = (K + 1) * 2.0 ** -24 * (au[rows].double().abs() @ bu.double().abs()) + 1e-300
    ratio = float(((outs[0][rows].double() - ex).abs() / bound).max().item())
    cu = torch.zeros((M, N), device="cuda")
    ob.mtm(cu, au, bu, None, variant="3xtf32", config=cfg, split_k=1)()
    torch.cuda.synchronize()
    res.append({"shape": shape, "cfg": cfg, "name": name, "exact": exact, "deterministic": same, "ratio": ratio,
                "explicit_split1_name": ob.last_choice()["name"]})
print(json.dumps(res))
"""


def test_stream_k_exact_and_deterministic(ob):
    """Opt-in stream-K scheduling (B200_TF32_STREAM_K=1, read once per process: runs in a child): the k-block iterations
    are cut into equal ranges per CTA group; tiles finished by several groups are added into C in a fixed order.
    Integer data: exact against fp64; uniform data: identical bits from call to call and within the tolerance; an
    explicit split factor of 1 switches it off."""
    import json
    import os
    import sys
    env = dict(os.environ, B200_TF32_STREAM_K="1")
    r = subprocess.run([sys.executable, "-c", _STREAM_K_CHILD % str(ROOT)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-2500:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("[")][-1])
    used = 0
    for e in res:
        print(f"\n[stream-K] {e['shape']} cfg {e['cfg']} -> {e['name']}: ratio {e['ratio']:.4f}")
        assert e["exact"] and e["deterministic"] and e["ratio"] <= TOL_C["3xtf32"], e
        assert "streamk" not in e["explicit_split1_name"], e
        used += "streamk" in e["name"]
    assert used >= 4, res


@pytest.mark.parametrize("dtype,variant", [(np.float32, "3xtf32"), (np.float32, "simt"), (np.float64, "dmma")])
def test_two_streams_share_the_workspaces_safely(dtype, variant, ob):
    """the asynchronous *_dev entries keep one grow-only workspace per device (lo planes, packed
    operands) for ALL streams.  Calls issued on two streams without any synchronisation in between — different
    problems, the second larger so the workspace is re-allocated while the first stream's kernels are still queued —
    must both come out exact: every use is ordered behind the previous one by an event, growth is stream-ordered."""
    import torch
    skip_if_absent(ob, dtype, variant)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    g = torch.Generator(device="cuda").manual_seed(5)
    mk = lambda r, c: torch.randint(0, 10, (r, c), device="cuda", generator=g).to(tdt)
    # odd leading dimensions on the first problem: its operands are packed into the shared workspace
    a1, b1, c1 = mk(1100, 1030)[:, :1027], mk(1027, 1210)[:, :1203], torch.zeros((1100, 1203), device="cuda", dtype=tdt)
    a2, b2, c2 = mk(2304, 2048), mk(2048, 2560), torch.zeros((2304, 2560), device="cuda", dtype=tdt)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    f1 = ob.mtm(c1, a1, b1, None, variant=variant, stream=s1.cuda_stream)
    f2 = ob.mtm(c2, a2, b2, None, variant=variant, stream=s2.cuda_stream)
    reps = 8
    for _ in range(reps):
        f1()
        f2()
    torch.cuda.synchronize()
    assert torch.equal(c1.double(), reps * (a1.double() @ b1.double())), f"{variant}: stream 1 corrupted"
    assert torch.equal(c2.double(), reps * (a2.double() @ b2.double())), f"{variant}: stream 2 corrupted"


@pytest.mark.parametrize("shape,cfg", [((2560, 2560, 2560), None), ((4096, 4096, 4096), 0), ((4096, 4096, 4096), 2), ((2560, 2816, 2048), 0),
                                       ((3328, 4096, 1024), 1)])
def test_tail_split_exact_and_deterministic(shape, cfg, ob):
    """More tiles than CTA groups and a ragged last wave: only that wave's tiles are cut along K (the other waves run
    whole tiles), their halves added into C in a fixed order.  Integer data exact, repeated calls bit-identical,
    tolerance kept; an explicit split factor of 1 switches it off."""
    import torch
    skip_if_absent(ob, np.float32, "3xtf32")
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randint(0, 10, (M, K), device="cuda", generator=g).float()
    b = torch.randint(0, 10, (K, N), device="cuda", generator=g).float()
    c0 = torch.randint(0, 10, (M, N), device="cuda", generator=g).float()
    c = c0.clone()
    fn = ob.mtm(c, a, b, None, variant="3xtf32", config=cfg)
    fn()
    fn()
    torch.cuda.synchronize()
    name = ob.last_choice()["name"]
    assert torch.equal(c.double(), c0.double() + 2 * (a.double() @ b.double())), name
    au = torch.rand((M, K), device="cuda", generator=g) * 2 - 1
    bu = torch.rand((K, N), device="cuda", generator=g) * 2 - 1
    outs = []
    for _ in range(3):
        cu = torch.zeros((M, N), device="cuda")
        ob.mtm(cu, au, bu, None, variant="3xtf32", config=cfg)()
        torch.cuda.synchronize()
        outs.append(cu)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), f"{name}: run-to-run differences"
    rows = torch.arange(0, M, max(1, M // 64), device="cuda")
    exact = au[rows].double() @ bu.double()
    bound = (K + 1) * 2.0 ** -24 * (au[rows].double().abs() @ bu.double().abs()) + 1e-300
    ratio = float(((outs[0][rows].double() - exact).abs() / bound).max().item())
    print(f"\n[tail split] {shape} cfg {cfg} -> {name}: ratio {ratio:.4f}")
    assert ratio <= TOL_C["3xtf32"]
    if shape != (3328, 4096, 1024):
        assert "tailsplit" in name, name
    cu = torch.zeros((M, N), device="cuda")
    ob.mtm(cu, au, bu, None, variant="3xtf32", config=cfg, split_k=1)()
    torch.cuda.synchronize()
    assert "split" not in ob.last_choice()["name"]
