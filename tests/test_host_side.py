"""CPU tests of the product's host side: the device FFT's index arithmetic (sequential emulation),
the C-ABI libraries' exported symbols, the host grain schedule against the oracle, and the loud
failure when no GPU is present."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(Path(__file__).resolve().parent))
import signals as S  # noqa: E402


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


SAN = ["-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-g"]


def _build_emulation(cmd):
    """The emulations index std::vectors exactly as the kernels index shared memory, so they are built
    with ASan + UBSan when the toolchain has them: an out-of-range slot becomes a test failure."""
    r = subprocess.run(cmd[:1] + SAN + cmd[1:], capture_output=True, text=True)
    if r.returncode != 0:
        subprocess.run(cmd, check=True, capture_output=True)


def test_device_fft_emulated_on_host(tmp_path):
    """melonix_b200/csrc/fft.cuh compiled by g++ with a sequential thread-group emulation."""
    exe = tmp_path / "fft_emul"
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    _build_emulation([gxx, "-std=c++17", "-O2", "-x", "c++", str(ROOT / "tests/host/fft_emul.cpp"), "-x", "c",
                      str(ROOT / "oracle/fft64.c"), "-o", str(exe), "-lm"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout


def test_spec_frame_steps_emulated_on_host(tmp_path):
    """Per-thread steps of the regular-hop Spec kernel (spec_frame.cuh) under a sequential thread-group
    emulation, against the oracle restatement of spec.cpp:44-66."""
    exe = tmp_path / "spec_frame_emul"
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    objs = []
    for c in ("spec_ref.c", "fft64.c"):
        o = tmp_path / (c + ".o")
        subprocess.run(["gcc", "-O2", "-fopenmp", "-c", str(ROOT / "oracle" / c), "-o", str(o)], check=True,
                       capture_output=True)
        objs.append(str(o))
    _build_emulation([gxx, "-std=c++17", "-O2", str(ROOT / "tests/host/spec_frame_emul.cpp"), *objs, "-o", str(exe),
                      "-lm", "-fopenmp"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout


def test_grain_segmentation_bit_logic_emulated_on_host(tmp_path):
    """Bit logic of the device grain segmentation (grain_seg.cuh) run sequentially against the oracle
    restatement of App::preproc (app.cpp:156-235)."""
    exe = tmp_path / "grain_seg_emul"
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    o = tmp_path / "grain_ref.o"
    subprocess.run(["gcc", "-O2", "-c", str(ROOT / "oracle/grain_ref.c"), "-o", str(o)], check=True, capture_output=True)
    _build_emulation([gxx, "-std=c++17", "-O2", str(ROOT / "tests/host/grain_seg_emul.cpp"), str(o), "-o", str(exe), "-lm"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout


def test_bin_shift_ring_arithmetic_on_host(tmp_path):
    """pv_shift.cuh (the phase-increment arithmetic of the analysis kernel, compiled here by g++): 4.4 M
    cases incl. INT_MIN advances, cut flips and empty K_j against a 128-bit evaluation of the defining
    formula (exact) and against the spec's double formula (within the one rounding both perform)."""
    exe = tmp_path / "pv_shift_emul"
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    _build_emulation([gxx, "-std=c++17", "-O2", str(ROOT / "tests/host/pv_shift_emul.cpp"), "-o", str(exe), "-lm"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout


def test_analysis_bin_cut_decisions_on_host(tmp_path):
    """pv_analysis.cuh (magnitude, integer-turn phase, wrapped advance and the FP64 decision at the +-pi cut,
    compiled here by g++) against the oracle's evaluation of PV-spec A.3 on 256 k bin-frames, half of them
    steered to within 1e-3 ... 1e-15 rad of the cut: never on the other side, values within 1e-6 rad, the
    integer phases telescope exactly."""
    exe = tmp_path / "pv_analysis_emul"
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    _build_emulation([gxx, "-std=c++17", "-O2", "-ffp-contract=off", str(ROOT / "tests/host/pv_analysis_emul.cpp"),
                      "-o", str(exe), "-lm"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout


def test_c_abi_exports_every_declared_symbol():
    from melonix_b200 import capi, hostlib
    L = capi.lib()
    syms = capi.declared_symbols()
    assert len(syms) >= 20
    assert [s for s in syms if not hasattr(L, s)] == []
    H = hostlib.lib()
    hs = hostlib.declared_symbols()
    assert len(hs) == 10 and [s for s in hs if not hasattr(H, s)] == []


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import melonix_b200 as m
    with pytest.raises(m.MlxError) as e:
        m.Engine(0)
    assert "no CPU fallback" in str(e.value)


def test_host_grain_schedule_matches_oracle(oracle):
    from melonix_b200 import hostlib as H
    x = S.two_tone(6.0)
    gs, gl = H.grain_segment(x)
    ogs, ogl = oracle.grain_segment(x)
    assert np.array_equal(gs, ogs) and np.array_equal(gl, ogl)
    for mk in ([], [(10, 0, 0, 3.0), (x.size - 10, 0, 0, 3.0)], [(10, 0, 0, -7.0), (x.size - 10, 0, 0, -7.0)],
               [(40000, 0, 0.5, 2.0), (120000, 0, -0.3, -1.5), (250000, 0, 0.0, 4.0)]):
        s = H.export_schedule(x, 48000, mk, gs, gl)
        o = oracle.grain_export(x, 48000, mk, ogs, ogl)
        for k in ("gstart", "glen", "rate", "next"):
            assert np.array_equal(s[k], o["schedule"][k]), k
        assert np.array_equal(s["out_off"][:-1], o["schedule"]["out_off"])
        assert s["out_off"][-1] + s["tail_zeros"] == o["pcm"].size
        for t in (0.0, 0.37, 1.9, 5.5, 7.0):
            assert H.time2sample(mk, 48000, t) == oracle.time2sample(mk, 48000, t)
            assert H.time2pitchbend(mk, 48000, x.size, t) == oracle.time2pitchbend(mk, 48000, x.size, t)
        for smp in (0, 1000, 100000, 260000):
            assert H.sample2time(mk, 48000, smp) == oracle.sample2time(mk, 48000, smp)


def test_grain_edge_cases(oracle):
    from melonix_b200 import hostlib as H
    for x in (np.zeros(0, np.float32), np.zeros(1200, np.float32), np.ones(5000, np.float32),
              S.two_tone(0.05), -S.two_tone(0.2)):
        gs, gl = H.grain_segment(x)
        ogs, ogl = oracle.grain_segment(x)
        assert np.array_equal(gs, ogs) and np.array_equal(gl, ogl)
        s = H.export_schedule(x, 48000, [], gs, gl)
        o = oracle.grain_export(x, 48000, [], ogs, ogl)
        assert s["out_off"][-1] + s["tail_zeros"] == o["pcm"].size


def test_bench_reference_arm_contract():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_picks_layout_is_a_pure_host_function(oracle):
    """mlx_picks_levels / mlx_picks_layout need no GPU: level count and offsets of the min/max pyramid
    (reference app.cpp:352-369) agree with the oracle for every size class."""
    import ctypes as C

    import numpy as np

    from melonix_b200 import capi
    L = capi.lib()
    for n in [0, 1, 2, 3, 4, 5, 7, 8, 9, 4095, 4096, 4097, 14_400_000, 345_600_000, 2**31 - 65537]:
        levels = L.mlx_picks_levels(n)
        off = np.zeros(levels + 1, np.int64)
        total = L.mlx_picks_layout(n, off.ctypes.data_as(C.c_void_p))
        o = oracle.lib()
        assert levels == o.mlxo_picks_levels(C.c_int64(n))
        ooff = np.zeros(levels + 1, np.int64)
        o.mlxo_picks_layout.restype = C.c_int64
        assert total == o.mlxo_picks_layout(C.c_int64(n), ooff.ctypes.data_as(C.c_void_p))
        assert np.array_equal(off, ooff)
        assert total == sum(n >> (l + 1) for l in range(levels))


def test_host_picks_queries_match_oracle(oracle):
    """melonix::Picks (host/picks.cpp): layout and single range queries on a pyramid, against the oracle's
    restatement of app.cpp:347-426 -- bit patterns, NaN and signed zeros included."""
    from melonix_b200 import hostlib as H
    rng = np.random.default_rng(11)
    for n in (0, 1, 2, 3, 9, 1000, 4097, 100003):
        x = rng.standard_normal(n).astype(np.float32)
        if n > 50:
            x[::7] = 0.0
            x[3::7] = -0.0
            x[5::101] = np.nan
        pairs, off = oracle.picks_build(x)
        assert np.array_equal(H.picks_layout(n), off)
        if n == 0:
            r = np.array([[0, 0], [0, 1], [-1, 3], [2, 1]], np.int32)
        else:
            s = rng.integers(0, max(n - 1, 1), 3000)
            e = np.minimum(s + (2.0 ** rng.uniform(0, np.log2(max(n, 2)), 3000)).astype(np.int64), n - 1)
            r = np.concatenate([np.stack([s, e], 1), [[5, 5], [7, 3], [n, n], [-3, 10], [10, -3], [0, n], [0, n - 1]]])
        r = r.astype(np.int32)
        got = H.minmax_ranges(x, pairs if pairs.size else np.zeros((1, 2), np.float32), r)
        exp = oracle.minmax_ranges(x, pairs, off, r)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), n


@pytest.mark.skipif(not Path("/root/reference/app.cpp").exists(), reason="needs the reference tree")
def test_reference_front_end_links_against_the_drop_in_classes(tmp_path):
    """INTEGRATION.md section 1, executed: a build directory with the reference's front-end sources
    (app.cpp, app.hpp, marker.hpp, save-wav.*, the dialog headers: symlinks, nothing copied into the repo)
    and OUR spec.hpp / spec.cpp / spec-cache.hpp / spec-cache.cpp / range.hpp / texture.hpp in place of the
    reference's.  The reference's app.cpp compiles unmodified against our class surface and links with
    libmelonix_b200.so (UI / audio / codec headers: the no-op shims of oracle/shim_app).  Without a GPU
    the resulting program stops where it must: Spec's constructor refuses to run on the CPU."""
    ref, host = Path("/root/reference"), ROOT / "melonix_b200" / "host"
    for f in ("app.cpp", "app.hpp", "file-open.hpp", "file-save-as.hpp", "marker.hpp", "save-wav.cpp", "save-wav.hpp"):
        (tmp_path / f).symlink_to(ref / f)
    for f in ("spec.hpp", "spec.cpp", "spec-cache.hpp", "spec-cache.cpp", "range.hpp", "texture.hpp", "colour_ramp.hpp"):
        (tmp_path / f).symlink_to(host / f)
    (tmp_path / "stubs.cpp").write_text(
        '#include "app.hpp"\n'
        "auto FileOpen::draw() -> bool { return false; }\n"
        "auto FileOpen::getSelectedFile() const -> std::filesystem::path { return {}; }\n"
        "FileSaveAs::FileSaveAs(std::string name) : dialogName(std::move(name)) { fileName.fill(0); }\n"
        "auto FileSaveAs::draw() -> bool { return false; }\n"
        "auto FileSaveAs::getSelectedFile() const -> std::string { return {}; }\n"
        'int main() { App app; app.openFile("/nonexistent.wav"); return 0; }\n')
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    flags = ["-std=c++20", "-O1", "-I", str(ROOT / "oracle/shim_app"), "-I", str(ROOT / "oracle/shim"),
             "-I", str(ROOT / "include"), "-I", str(tmp_path)]
    objs = []
    for f in ("app.cpp", "spec.cpp", "spec-cache.cpp", "save-wav.cpp", "stubs.cpp"):
        o = tmp_path / (f + ".o")
        r = subprocess.run([gxx, *flags, "-c", str(tmp_path / f), "-o", str(o)], capture_output=True, text=True)
        assert r.returncode == 0, f + "\n" + r.stderr[-3000:]
        objs.append(str(o))
    lib = ROOT / "melonix_b200"
    r = subprocess.run([gxx, "-o", str(tmp_path / "melonix_dropin"), *objs, "-L", str(lib), "-lmelonix_b200",
                        f"-Wl,-rpath,{lib}", "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    if not _has_gpu():
        r = subprocess.run([str(tmp_path / "melonix_dropin")], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_drop_in_build_refuses_to_run_without_a_gpu(oracle):
    """oracle/_ref/libapp_dropin.so (reference front-end + product Spec/SpecCache, oracle/Makefile `dropin`)
    loads next to the all-reference build and, without a B200, fails loudly in Spec's constructor."""
    if not oracle.have_dropin_app():
        pytest.skip("oracle/_ref/libapp_dropin.so not built (needs /root/reference)")
    if _has_gpu():
        pytest.skip("checks the no-GPU failure mode")
    x = S.two_tone(1.0)
    with oracle.RefApp(x, 48000, []) as ref:
        assert ref.grains()[0].size > 0
    with pytest.raises(RuntimeError) as e:
        oracle.RefApp(x, 48000, [], build="dropin")
    assert "no CPU fallback" in str(e.value)


def test_host_colour_ramp_equals_reference_ramp(oracle):
    """mlxh_colour_ramp (what the drop-in Spec recolours cached columns with on a brightness change)
    against the oracle's ramp, which is pinned byte for byte to the reference's populateTex
    (tests/test_oracle.py): all three segments, the thresholds, clamping, negative and NaN-free input."""
    from melonix_b200 import hostlib as H
    rng = np.random.default_rng(5)
    for k in (2.0 ** 15, 2.0 ** 11, 700.0, 1.0):
        v = np.concatenate([rng.random(20000).astype(np.float32) * np.float32(300.0 / k),
                            np.array([0, 84.999, 85, 85.001, 169.999, 170, 170.001, 254.999, 255, 256, -1], np.float32)
                            / np.float32(k)])
        assert np.array_equal(H.colour_ramp(v, k), oracle.colormap(v, float(k)))
