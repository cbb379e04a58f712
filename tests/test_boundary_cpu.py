"""CPU-side checks of the drop-in boundary (no compute without a GPU):
the C-ABI library loads, exports every symbol include/aerobulk_gpu.h declares, its pure-host entry
points behave, compute entry points fail loudly without a device, and the C++ API links."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "aerobulk_gpu.h")


@pytest.fixture(scope="module")
def ab():
    from aerobulk_b200 import build
    build.build()
    import aerobulk_b200 as ab
    ab.lib()
    return ab


def _declared_symbols():
    txt = open(HDR).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aerobulk_(?:gpu_\w+|cxx_skin|cxx_no_skin))\s*\(", txt)))


def test_library_exports_every_declared_symbol(ab):
    L = ab.lib()
    syms = _declared_symbols()
    assert "aerobulk_cxx_skin" in syms and "aerobulk_cxx_no_skin" in syms and len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} is declared in include/aerobulk_gpu.h but not exported"


def test_header_is_plain_c_and_cxx():
    """include/aerobulk_gpu.h is what a C (cgo, ctypes generators ...) or C++ caller includes: strict C99 and C++11."""
    for comp, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "c++")):
        r = subprocess.run([comp, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", lang, HDR],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_no_unexpected_exports(ab):
    """Only the declared C ABI and the aerobulk:: C++ API are exported (-fvisibility=hidden)."""
    from aerobulk_b200.model import _SO
    out = subprocess.check_output(["nm", "-D", "--defined-only", _SO], text=True)
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    declared = set(_declared_symbols())
    for n in names:
        assert n in declared or n.startswith("_ZN8aerobulk"), n


def test_pure_host_entry_points(ab):
    L = ab.lib()
    assert ab.work_per_point("ncar", False, 5) == 1947 + 5 * 595
    assert ab.work_per_point("coare3p6", True, 5) == 4698 + 5 * 4733
    assert ab.bytes_per_point("ncar", False) == 88 and ab.bytes_per_point("coare3p0", True) == 176
    assert ab.bytes_per_point("ecmwf", True) == 128 and ab.bytes_per_point("andreas", True) == 88
    ab.reset()
    assert ab.nb_iter() == 5 and not ab.use_skin() and ab.humidity_type() == "sh"
    ab.set_nb_iter(7)
    assert ab.nb_iter() == 7
    ab.reset()
    assert ab.nb_iter() == 5
    assert list(ab.diag_reduce_ops()) == [0] + [0, 1, 2] * 6          # count, then sum / min / max of the six flux fields
    ops = [L.aerobulk_gpu_stats_reduce_op(i) for i in range(64)]
    assert ops[:2] == [0, 0] and ops[2:7] == [0, 1, 2, 1, 2] and ops[47:] == [0] * 17


@pytest.mark.skipif("torch" in sys.modules and __import__("torch").cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_compute_fails_loudly_without_device(ab):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    x = np.full(4, 290.0)
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, x, x, x * 0 + 0.01, x * 0 + 5, x * 0, x * 0 + 101000.0)
    assert e.value.code == 100 and "no CPU fallback" in e.value.message
    # every compute entry point, not only aerobulk_model
    from aerobulk_b200 import synth
    d = synth.station_series(3, 2)
    one = np.ones(3)
    for call in (lambda: ab.series("ncar", 2.0, 10.0, **d),
                 lambda: ab.oce_ice("nemo", None, 2.0, 10.0, **synth.ice_fields(4)),
                 lambda: ab.turb_ice("nemo", 2.0, 10.0, one * 270, one * 271, one * 1e-3, one * 1e-3, one * 3),
                 lambda: ab.series_ice("nemo", 2.0, 10.0, one * 0.9, one * 265, one * 266, one * 1e-3, one * 4, one * 1e5, one * 50, one * 200),
                 lambda: ab.turb("ncar", 1, 2.0, 10.0, one * 290, one * 289, one * 1e-2, one * 8e-3, one * 5),
                 lambda: ab.flux_diagnostics({"QL": one})):
        with pytest.raises(ab.AerobulkError) as e:
            call()
        assert e.value.code == 100
    # argument errors are detected before any device work
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(0, 1, "ncar", 2.0, 10.0, x, x, x, x, x, x)
    assert e.value.code == 1


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under aerobulk_b200/ may reference it."""
    pkg = os.path.join(ROOT, "aerobulk_b200")
    for base, _, files in os.walk(pkg):
        if os.path.basename(base) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".f90")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), (base, f)


def test_cpp_api_links_and_reports(ab, tmp_path):
    """The C++ API (include/aerobulk.hpp) compiles and links against libaerobulk_gpu.so like the
    reference's `make cpp` target (Makefile:120-122) links -laerobulk_cxx -laerobulk."""
    from aerobulk_b200.model import _SO
    exe = str(tmp_path / "example_call_aerobulk_cxx.x")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "example_call_aerobulk.cpp"),
                           "-o", exe, "-L", os.path.dirname(_SO), "-laerobulk_gpu",
                           "-Wl,-rpath," + os.path.dirname(_SO)])
    r = subprocess.run([exe], capture_output=True, text=True)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except ImportError:
        has_gpu = False
    if has_gpu:
        assert r.returncode == 0 and r.stdout.count("RESULT") == 5, r.stdout + r.stderr
    else:
        # fail-stop like the reference's ctl_stop: message on stdout, process ends (exit status 1)
        assert r.returncode == 1 and "E R R O R" in r.stdout and "no CUDA device" in r.stdout, r.stdout + r.stderr


def test_series_cli_arguments():
    """Argument handling of the series CLI needs no GPU (it mirrors the prompts of the reference program)."""
    r = subprocess.run([sys.executable, "-m", "aerobulk_b200.series_cli", "a.csv", "b.csv", "--zu", "150"], cwd=ROOT,
                       capture_output=True, text=True)
    assert r.returncode == 2 and "Be reasonable" in r.stderr
    r = subprocess.run([sys.executable, "-m", "aerobulk_b200.series_cli", "a.csv", "b.csv", "--algo", "coare9"], cwd=ROOT,
                       capture_output=True, text=True)
    assert r.returncode == 2 and "invalid choice" in r.stderr
    # ocean and sea-ice algorithms do not mix
    for args, word in ((["--ice", "--algo", "ncar"], "does not go with"), (["--algo", "lg15"], "needs --ice")):
        r = subprocess.run([sys.executable, "-m", "aerobulk_b200.series_cli", "a.csv", "b.csv"] + args, cwd=ROOT,
                           capture_output=True, text=True)
        assert r.returncode == 2 and word in r.stderr, r.stderr


def test_host_copy_threads_selftest(ab):
    """The copy threads behind the pageable-array path move data faithfully (many back-to-back jobs, odd sizes)."""
    L = ab.lib()
    L.aerobulk_gpu_selftest_host_copy.restype = C.c_int
    L.aerobulk_gpu_selftest_host_copy.argtypes = [C.c_longlong, C.c_int]
    assert L.aerobulk_gpu_selftest_host_copy(1, 3) == 0
    assert L.aerobulk_gpu_selftest_host_copy(16384 * 37 + 5, 50) == 0
    assert L.aerobulk_gpu_selftest_host_copy(1_000_003, 20) == 0


def test_host_copy_pool_under_thread_sanitizer(tmp_path):
    """aerobulk_b200/csrc/ab_copy_pool.hpp built with -fsanitize=thread: 400 back-to-back jobs, no race, no lost piece."""
    exe = str(tmp_path / "cp_tsan")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread",
                        os.path.join(ROOT, "tests", "copy_pool_tsan.cpp"), "-o", exe], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime here: " + r.stderr[-200:])
    for args in ([], ["streaming-stores"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        if "unexpected memory mapping" in r.stderr:   # the sanitizer runtime cannot start under this kernel's ASLR settings
            pytest.skip("ThreadSanitizer runtime cannot run here")
        assert r.returncode == 0 and "bad=0" in r.stdout and "ThreadSanitizer" not in r.stderr, r.stderr[-2000:]


def test_chunk_plan_of_host_array_calls(ab):
    """The row-block plans of aerobulk_gpu_model (pure host logic): boundaries start at 0, end at n, never decrease, fall on
    whole 2048-point windows, stay within the chunk limits; the staged pipeline shrinks its chunks, the pageable path
    ramps them up and down."""
    L = ab.lib()
    L.aerobulk_gpu_chunk_plan.restype = C.c_int
    L.aerobulk_gpu_chunk_plan.argtypes = [C.c_longlong, C.c_int, C.POINTER(C.c_longlong)]

    def plan(n, kind):
        cs = (C.c_longlong * 17)()
        m = L.aerobulk_gpu_chunk_plan(n, kind, cs)
        return m, [cs[i] for i in range(m + 1)]

    rng = np.random.default_rng(0)
    sizes = [0, 1, 5, 2047, 2048, 2049, 65535, 65536, 199999, 200000, 400001, 1036800, 9331200, 83980800] + \
        [int(x) for x in rng.integers(1, 20_000_000, 40)]
    for n in sizes:
        for kind in (0, 1, 2, 3):
            m, cs = plan(n, kind)
            assert 1 <= m <= 16 and cs[0] == 0 and cs[-1] == n, (n, kind, cs)
            assert all(b >= a for a, b in zip(cs, cs[1:])), (n, kind, cs)
            assert all(c % 2048 == 0 for c in cs[:-1]), (n, kind, cs)
            if kind == 0:
                assert m == 1
    m, cs = plan(1036800, 1)
    sz = [b - a for a, b in zip(cs, cs[1:])]
    assert m == 5 and sz == sorted(sz, reverse=True)                 # K..1 weights: short exposed tail
    m, cs = plan(1036800, 2)
    sz = [b - a for a, b in zip(cs, cs[1:])]
    assert m == 6 and sz[0] < sz[1] < sz[2] and sz[3] > sz[4] > sz[5] and abs(sz[0] - sz[5]) <= 4096
    m, cs = plan(83980800, 3)                                        # jt == 1 with the speculative AEROBULK_INIT: both
    sz = [b - a for a, b in zip(cs, cs[1:])]                         # PCIe directions busy -> equal pieces, half-size ends
    assert m == 12 and max(sz[1:-1]) - min(sz[1:-1]) <= 4096 and sz[0] < 0.6 * sz[1] and sz[-1] < 0.6 * sz[1]
    assert L.aerobulk_gpu_chunk_plan(-1, 0, (C.c_longlong * 17)()) == -1 and L.aerobulk_gpu_chunk_plan(10, 4, (C.c_longlong * 17)()) == -1


def test_python_mirror_normalises_layout_and_validates_out(ab):
    """ADVICE r1 (model.py:127): every 2-D field reaches the library column-major, whatever layout the caller used, and a
    caller-supplied output array of another dtype / shape / layout is refused instead of being written with a flat index
    that pairs different grid points across fields."""
    from aerobulk_b200 import model
    c = np.arange(12, dtype=np.float64).reshape(3, 4)                 # numpy default: C order
    f = model._f64(c)
    assert f.flags.f_contiguous and np.array_equal(f, c)
    assert np.array_equal(f.ravel(order="K"), c.ravel(order="F"))     # memory order == the reference's ji + Ni*jj
    fo = np.asfortranarray(c)
    assert model._f64(fo) is fo                                       # already right: no copy (pinned slabs stay pinned)
    v = np.arange(10.0)[::2]
    assert model._f64(v).flags.c_contiguous
    assert model._f64(np.arange(6, dtype=np.float32).reshape(2, 3)).dtype == np.float64
    with pytest.raises(ab.AerobulkError):
        model._f64(c, (4, 3))
    good = np.empty((3, 4), order="F")
    assert model._check_out("QL", good, (3, 4)) is good
    for bad in (np.empty((3, 4)), np.empty((3, 4), dtype=np.float32, order="F"), np.empty((4, 3), order="F"), [0.0] * 12,
                np.empty(24)[::2]):
        with pytest.raises(ab.AerobulkError):
            model._check_out("QL", bad, (3, 4) if not (isinstance(bad, np.ndarray) and bad.ndim == 1) else (12,))
    ro = np.empty((3, 4), order="F")
    ro.flags.writeable = False
    with pytest.raises(ab.AerobulkError):
        model._check_out("QL", ro, (3, 4))
    # the public entry validates before anything is computed (no device needed to be refused)
    z = np.full((3, 4), 290.0)
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, z, z, z * 0 + 0.01, z * 0 + 5, z * 0, z * 0 + 101000.0, out={"QL": np.empty((3, 4))})
    assert e.value.code == 101


def test_shard_plan_of_split_calls(ab):
    """aerobulk_gpu_set_devices(n): contiguous shards covering [0, n), inner boundaries on multiples of 2048 points (whole
    sort windows / thread blocks), never more shards than devices, fewer for tiny grids (needs no device)."""
    for n, nd in ((1036800, 8), (83980800, 8), (64800, 2), (5000, 8), (1, 4), (2048 * 3 + 1, 3)):
        plan = ab.shard_plan(n, nd)
        k = len(plan) - 1
        assert 1 <= k <= nd and plan[0] == 0 and plan[-1] == n
        assert all(b > a for a, b in zip(plan[:-1], plan[1:]))
        assert all(b % 2048 == 0 for b in plan[1:-1])
        sizes = [b - a for a, b in zip(plan[:-1], plan[1:])]
        assert max(sizes) - min(sizes[:-1] or sizes) <= 0 or max(sizes[:-1]) == min(sizes[:-1])   # equal shards but the last
        assert max(sizes) <= -(-n // k) + 2048
    assert ab.shard_plan(83980800, 8) == [i * 10498048 if i < 8 else 83980800 for i in range(9)]   # C5: 810 rows rounded up to 2048 points
    with pytest.raises(ValueError):
        ab.shard_plan(-1, 2)
    assert ab.get_devices() == 1


def test_design_tables_quote_the_committed_evidence():
    """The measured tables of DESIGN.md are exactly what tools/design_numbers.py generates from the evidence files
    profiles/CURRENT.json names (bench line, ncu counters, launch list, kbench) -- the document cannot drift from them."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "design_numbers.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
