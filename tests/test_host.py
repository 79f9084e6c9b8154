"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol the header declares,
the native grid front end (ugx reader + regular refinement) matches the oracle bit for bit, error behaviour,
and the world_size-2 plumbing on gloo.  No GPU, no compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import GRID2D, GRID3D, ROOT


def test_library_exports_every_declared_symbol():
    from admm_optim_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "admm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ab_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 60
    for n in sorted(names):
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    assert names - {"ab_last_error", "ab_version"} == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.ab_version() >= 100


def _host_domain(path_or_npz, refs):
    from admm_optim_b200 import ug4
    ug = ug4.Backend.host_only()
    dom = ug4.Domain(ug)
    ug.LoadDomain(dom, path_or_npz)
    ug.util.refinement.CreateRegularHierarchy(dom, refs, False, None)
    return dom


@pytest.mark.parametrize("grid,refs", [(GRID3D, 2), (GRID2D, 3)])
def test_native_refinement_is_bit_identical_to_oracle(grid, refs):
    from oracle import mesh_np as M
    dom = _host_domain(grid, refs)
    levels = M.build_hierarchy(M.load_npz(grid), refs)
    assert dom.num_levels() == refs + 1
    for l, o in enumerate(levels):
        g = dom.get_level(l)
        assert np.array_equal(g["xyz"], o.xyz) and np.array_equal(g["elems"], o.elems) and np.array_equal(g["vsub"], o.vsub)
        if l:
            assert np.array_equal(g["parent_a"], o.parent_a) and np.array_equal(g["parent_b"], o.parent_b)
        assert dom.level_info(l)["nedges"] == len(M.unique_edges(o))


def test_ugx_reader_native_vs_oracle(tmp_path):
    """A small hand-written .ugx (two tetrahedra, two subsets) through both readers; plus the shipped grids when
    the reference tree is present (it is not on the GPU box)."""
    from oracle import mesh_np as M
    ugx = textwrap.dedent("""\
        <?xml version="1.0" encoding="utf-8"?>
        <grid name="defGrid">
        <vertices coords="3">0 0 0 1 0 0 0 1 0 0 0 1 1 1 1</vertices>
        <edges>0 1 0 2 0 3 1 2 1 3 2 3 1 4 2 4 3 4</edges>
        <triangles>0 1 2 0 1 3 0 2 3 1 2 3 1 2 4 1 3 4 2 3 4</triangles>
        <tetrahedrons>0 1 2 3 1 2 3 4</tetrahedrons>
        <subset_handler name="defSH">
        <subset name="outer" color="0 0 0 1"><volumes>0 1</volumes><faces>3</faces><vertices>4</vertices><edges>6 7 8</edges></subset>
        <subset name="wall" color="1 0 0 1"><faces>0 1 2 4 5 6</faces><edges>0 1 2 3 4 5</edges><vertices>0 1 2 3</vertices></subset>
        </subset_handler>
        </grid>
        """)
    f = tmp_path / "two_tets.ugx"
    f.write_text(ugx)
    files = [str(f)] + [p for p in ("/root/reference/grids/refined.ugx", "/root/reference/grids/box_3D_elongated.ugx") if os.path.exists(p)]
    for path in files:
        dom = _host_domain(path, 1)
        levels = M.build_hierarchy(M.load_ugx(path), 1)
        for l, o in enumerate(levels):
            g = dom.get_level(l)
            assert np.array_equal(g["xyz"], o.xyz) and np.array_equal(g["elems"], o.elems) and np.array_equal(g["vsub"], o.vsub)
        for i, n in enumerate(levels[0].subset_names):
            assert dom.subset_index(n) == i


def test_shipped_fixtures_match_reference_grids_when_present():
    from oracle import mesh_np as M
    for name, npz in (("refined", GRID2D), ("box_3D_elongated", GRID3D)):
        src = "/root/reference/grids/%s.ugx" % name
        if not os.path.exists(src):
            pytest.skip("reference tree not present (GPU box)")
        a, b = M.load_ugx(src), M.load_npz(npz)
        assert np.array_equal(a.xyz, b.xyz) and np.array_equal(a.elems, b.elems) and np.array_equal(a.vsub, b.vsub)
        assert a.subset_names == b.subset_names and np.array_equal(a.sp_edges, b.sp_edges) and np.array_equal(a.sp_faces, b.sp_faces)


def test_error_behaviour_without_gpu_is_loud():
    from admm_optim_b200 import _lib, ug4
    lib = _lib.load()
    with pytest.raises(_lib.AdmmB200Error):
        _lib.call("ab_domain_load_ugx", None, b"/nonexistent/file.ugx", C.byref(C.c_void_p()))
    assert b"cannot open" in lib.ab_last_error()
    dom = _host_domain(GRID3D, 0)
    sp = C.c_void_p()
    with pytest.raises(_lib.AdmmB200Error, match="context"):      # spaces need a GPU context: no CPU fallback
        _lib.call("ab_space_create", dom.h, 1, 3, C.byref(sp))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(_lib.AdmmB200Error, match="no CPU fallback|no CUDA"):
            ug4.Backend(device=0)


def test_driver_rejects_unknown_parameters_and_descriptor_mirrors_reference():
    from admm_optim_b200 import driver
    with pytest.raises(ValueError):
        driver.ObstacleOptim(object(), 3, nonsense=1)
    assert driver.DEFAULTS_3D["numRefs"] == 2 and driver.DEFAULTS_3D["admmSteps"] == 2      # 3d_admm.lua:46,48
    assert driver.DEFAULTS_2D["numRefs"] == 3 and driver.DEFAULTS_2D["admmSteps"] == 1000   # 2d_admm.lua:43,45

    class Rec:
        def SuperLU(self):
            return "superlu"

        class util:
            class solver:
                @staticmethod
                def CreateSolver(desc):
                    return desc
    d3 = driver.linear_solver(Rec(), "dd", "space", False, 3)
    d2 = driver.linear_solver(Rec(), "dd", "space", True, 2)
    assert d3["type"] == "bicgstab" and d3["precond"]["type"] == "gmg" and d3["precond"]["smoother"] == "gs"
    assert (d3["precond"]["preSmooth"], d3["precond"]["postSmooth"], d3["precond"]["baseLevel"], d3["precond"]["rap"]) == (3, 3, 0, True)
    assert (d3["convCheck"]["iterations"], d3["convCheck"]["absolute"]) == (3000, 1e-10)      # obstacle_optim_3d_util.lua:34-35
    assert (d2["convCheck"]["iterations"], d2["convCheck"]["absolute"]) == (2000, 1e-12)      # obstacle_optim_util.lua:35-36


def test_bench_reference_arm_runs_under_gloo_world_size_2(tmp_path):
    """N>1 contract of the reference arm: rank 0 alone prints the line, other ranks exit 0 without work.
    Uses a tiny refinement so that the CPU suite stays fast."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    outs = []
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                                       "--warmup", "0", "--refs", "0"], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0, err
        outs.append(out.strip())
    import json
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert outs[1] == ""
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.common_config(0)          # the driver compares the two arms' `config`: one object for both


_GERSH_EMU = {}


def _gershgorin_emulation_library():
    """tests/cuda_host_shim/gershgorin_emulation.cpp (the library's exact-Gershgorin / pack / unpack kernel SOURCE compiled for the
    host) as a shared library, built once per session; None without a C++ compiler."""
    import shutil
    import tempfile
    if "path" not in _GERSH_EMU:
        _GERSH_EMU["path"] = None
        if shutil.which("g++"):
            out = os.path.join(tempfile.mkdtemp(prefix="ab_emu_"), "libgersh_emu.so")
            r = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-w", "-I" + os.path.join(ROOT, "admm_optim_b200", "csrc"),
                                os.path.join(ROOT, "tests", "cuda_host_shim", "gershgorin_emulation.cpp"), "-o", out], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-3000:]
            _GERSH_EMU["path"] = out
    return _GERSH_EMU["path"]


def _run_dist_host_workers(grid, refs, gather_dofs, port, world=2):
    import json
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), ADMM_B200_GATHER_DOFS=str(gather_dofs))
    if _gershgorin_emulation_library():
        env["ADMM_B200_GERSH_EMU"] = _gershgorin_emulation_library()
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_host_worker.py"), grid, str(refs)],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-2000:]
        outs.append(out)
    return json.loads(outs[0].strip().splitlines()[-1])


@pytest.mark.parametrize("grid,refs,counts,gather_dofs,lg", [(GRID3D, 2, [338, 2124, 14910], 7000, 1), (GRID3D, 2, [338, 2124, 14910], 100, 0),
                                                            (GRID2D, 2, [160, 596, 2296], 1500, 1)])
def test_partition_and_interfaces_gloo_world_size_2(grid, refs, counts, gather_dofs, lg):
    """Host side of the multi-GPU path on 2 gloo ranks: every global vertex is owned exactly once on every level,
    both sides of an interface list the same coordinates in the same order; the levels up to the gather level have exact
    local -> global vertex and matrix-block maps (the per-rank element-incidence counts add up to the global ones)."""
    res = _run_dist_host_workers(grid, refs, gather_dofs, 29613)
    assert all(r["decomposed"] and r["gather_level"] == lg for r in res)
    assert res[0]["blocks_ok"]
    # exact Gershgorin rows through the shared-block lists (NumPy twin of the device steps) = rows of the global operator
    assert res[0]["gershgorin"]["ok"] and res[0]["gershgorin"]["shared_blocks"] > 0, res[0]["gershgorin"]
    if _gershgorin_emulation_library():      # the kernel source of the library, compiled for the host, gave the global row sums on every rank
        emu = res[0]["gershgorin"]["kernel_source_emulation_worst"]
        assert len(emu) == 2 and all(e is not None and e < 1e-12 for e in emu), emu
    lv = [r["levels"] for r in res]
    for level in range(refs + 1):
        assert sum(r[level]["owned"] for r in lv) == counts[level]
        a, b = lv[0][level]["shared"].get("1"), lv[1][level]["shared"].get("0")
        assert a is not None and a == b and len(a) > 0
        assert lv[0][level]["nv"] + lv[1][level]["nv"] - len(a) == counts[level]


def test_partition_with_vertices_shared_by_four_ranks_gloo_world_size_4():
    """Four ranks on the 3D box: every rank neighbours every other one (edges shared by all four).  Ownership is unique, the
    vertical-interface maps are exact, and the shared-block lists reproduce the global row sums although blocks are held by up to
    four ranks."""
    res = _run_dist_host_workers(GRID3D, 2, 7000, 29615, world=4)
    assert all(r["decomposed"] and r["gather_level"] == 1 for r in res) and res[0]["blocks_ok"]
    g = res[0]["gershgorin"]
    assert g["ok"] and g["rows_where_loose_differs"] > 0 and g["shared_blocks"] > 0, g
    if _gershgorin_emulation_library():
        assert len(g["kernel_source_emulation_worst"]) == 4 and all(e is not None and e < 1e-12 for e in g["kernel_source_emulation_worst"]), g
    lv = [r["levels"] for r in res]
    for level, count in enumerate([338, 2124, 14910]):
        assert sum(r[level]["owned"] for r in lv) == count
    assert all(len(r[2]["shared"]) == 3 for r in lv)              # three neighbours each on the top level


def test_small_problem_is_not_decomposed_gloo_world_size_2():
    """A hierarchy whose top level is below the agglomeration threshold runs undivided: every rank keeps the whole grid
    (3d_admm.lua's default refinement has 44 730 unknowns -- decomposing it only adds latency)."""
    res = _run_dist_host_workers(GRID3D, 2, 400000, 29614)
    assert all(not r["decomposed"] and r["levels"] == [338, 2124, 14910] for r in res)


def test_agglomeration_level_choice():
    from admm_optim_b200 import partition as P
    z = np.load(GRID3D)
    g = {k: z[k] for k in z.files}
    nv = P.global_level_counts(g, 6)
    assert nv == [338, 2124, 14910, 111386, 860338, 6761314, 53608130]
    assert P.gather_level(nv, 3, 400000) == 3 and P.gather_level(nv[:3], 3, 400000) == 2 and P.gather_level(nv, 3, 10) == 0
    z2 = np.load(GRID2D)
    nv2 = P.global_level_counts({k: z2[k] for k in z2.files}, 7)
    assert nv2[3] == 9008 and nv2[7] == 2263808 and P.gather_level(nv2, 2, 400000) == 5


def test_rcb_partition_is_balanced_and_deterministic():
    from admm_optim_b200 import partition as P
    z = np.load(GRID3D)
    cent = z["xyz"][z["elems"]].mean(axis=1)
    for n in (2, 3, 4, 8):
        part = P.rcb_partition(cent, n)
        cnt = np.bincount(part, minlength=n)
        assert cnt.min() >= len(cent) // n - 1 and cnt.max() <= len(cent) // n + n
        assert np.array_equal(part, P.rcb_partition(cent, n))
        sub = P.extract_submesh({k: z[k] for k in z.files}, part, 0)
        assert sub["elems"].max() == len(sub["l2g"]) - 1 and (np.diff(sub["l2g"]) > 0).all()


def test_pdl_kernels_wait_before_touching_chain_data():
    """Static SASS audit of the programmatic-dependent-launch chain ("Latency path"): every kernel that is
    launched with AB_LAUNCH_PDL must contain griddepcontrol.wait (SASS ACQBULK), and nothing may be stored or reduced to global
    memory before the first wait.  Loads before the wait are allowed for data that is static during a solve; the audit tool
    (tools/check_pdl_sass.py) lists them."""
    import re
    import shutil
    import subprocess
    from conftest import ROOT
    if not shutil.which("cuobjdump") or not shutil.which("cu++filt"):
        pytest.skip("CUDA binary utilities not installed")
    lib = os.path.join(ROOT, "admm_optim_b200", "libadmm_b200.so")
    src = open(os.path.join(ROOT, "admm_optim_b200", "csrc", "lib.cu")).read()
    launched = set(re.findall(r"AB_LAUNCH_PDL\(ctx,\s*\(?\s*(k_\w+)", src))
    assert {"k_bsr_spmv_tma", "k_restrict", "k_prolong_add", "k_coarse_solve", "k_bicg_xr", "k_bicg_fused_first"} <= launched
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    seen = set()
    for block in re.split(r"\n\s*Function : ", sass)[1:]:
        mangled = block.split("\n", 1)[0].strip()
        name = subprocess.run(["cu++filt", mangled], capture_output=True, text=True).stdout
        m = re.search(r"ab::(k_\w+)", name)
        if not m or m.group(1) not in launched:
            continue
        seen.add(m.group(1))
        lines = block.split("\n")
        waits = [i for i, l in enumerate(lines) if "ACQBULK" in l]
        assert waits, "%s is launched with the PDL attribute but never waits" % m.group(1)
        early = [l for l in lines[:waits[0]] if re.search(r"\b(STG|ATOMG|RED)\b", l)]
        assert not early, "%s writes global memory before griddepcontrol.wait: %s" % (m.group(1), early[:2])
    assert launched <= seen, launched - seen
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_pdl_sass.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]


def test_bench_algorithmic_byte_formulas_match_survey():
    """bench.py computes `roofline.achieved` from the ALGORITHMIC bytes of SURVEY.md 8(d): B_spmv = nnzb(8 d^2 + 4) + nb(4 + 16 d),
    B_V = sum_l [7 B_spmv(A_l) + 2 B_P(l)], with (nb, nnzb) = (V, V + 2E) from the refinement recurrences.  Pin the level
    counts and the byte figures quoted in SURVEY / DESIGN (3D: L2 16.6 MB, L4 1.007 GB, L5 7.99 GB; V-cycle L4 8.18 GB,
    L5 64.8 GB; 2D L3 2.57 MB)."""
    sys.path.insert(0, ROOT)
    import bench
    lv = bench.global_counts(5, 3)
    assert [v for v, _ in lv] == [338, 2124, 14910, 111386, 860338, 6761314]
    assert lv[2] == (14910, 14910 + 2 * 96476) and lv[4][1] == 12662290 and lv[5][1] == 100454946
    assert abs(bench.spmv_bytes(3, *lv[2]) / 16.6e6 - 1) < 0.01
    assert bench.spmv_bytes(3, *lv[4]) == 1007071616
    assert bench.spmv_bytes(3, *lv[5]) == 7986164224
    assert bench.vcycle_bytes(3, lv[:5]) == 8183626352
    assert abs(bench.vcycle_bytes(3, lv) / 64.8e9 - 1) < 0.005
    # bytes the V(3,3) kernels really stream (6 matrix passes per level): within 10 % of the ncu DRAM sum of one cycle at numRefs 5
    # (62.03 GB read + 2.11 GB written, profiles/r02_vcycle_L5_dram_launches.csv) and below the 7-pass SURVEY figure
    assert abs(bench.vcycle_dram_bytes(3, lv) / 64.13e9 - 1) < 0.10 and bench.vcycle_dram_bytes(3, lv) < bench.vcycle_bytes(3, lv)
    l2 = bench.global_counts(3, 2)
    assert l2[3] == (9008, 9008 + 2 * 26672)
    assert abs(bench.spmv_bytes(2, *l2[3]) / 2.57e6 - 1) < 0.01


def test_gnuplot_write_data_restates_lua_semantics(tmp_path):
    """gnuplot.write_data as the scripts use it (3d_admm.lua:1274): columns = Lua tables, rows from index 1 up to the first gap,
    Lua 5.1 number formatting (%.14g), blank-separated."""
    from admm_optim_b200 import gnuplot
    f = str(tmp_path / "t.txt")
    step = {0: 0, 1: 1, 2: 2}                 # filled at index admm_steps = 0, 1, 2 like vADMM_Step
    val = {0: 0.5, 1: 1.0 / 3.0, 2: 1e-13}
    assert gnuplot.write_data(f, [step, val], False) == 2            # index 0 is not in the array part of a Lua table
    assert open(f).read() == "1 0.33333333333333 \n2 1e-13 \n"
    assert gnuplot.write_data(f, [step, val], False, first_row=0) == 3
    assert gnuplot.read_data(f) == [[0.0, 0.5], [1.0, 0.33333333333333], [2.0, 1e-13]]
    assert gnuplot.write_data(f, [[1, 2, 3], [4.0, 5.5]]) == 2       # the shortest column ends the table
    assert gnuplot.lua_number(3.0) == "3" and gnuplot.lua_number(0.1 + 0.2) == "0.3" and gnuplot.lua_number(123456789012345678.0) == "1.2345678901235e+17"


def test_trace_files_match_committed_golden(tmp_path):
    """The driver replay writes __ADMMStats_step_<k>_.txt / __NewtonStats_... / __NewtonIterations_... (3d_admm.lua:1274-1276,
    1307-1311) at the scripts' places; on the oracle backend the files must reproduce tests/golden/traces_3d_refs1
    (tools/make_golden.py traces): integer columns exactly, floating-point columns to 1e-9 relative."""
    from admm_optim_b200 import gnuplot
    from admm_optim_b200.driver import ObstacleOptim
    from oracle import ug4_np
    p = ObstacleOptim(ug4_np.Backend(smoother="cheb"), 3, numRefs=1, grid=GRID3D, admmSteps=3, trace_dir=str(tmp_path), newton_output=True).setup()
    p.set_sensitivity(p.synthetic_sensitivity(0.5))
    p.run_admm()
    gold = os.path.join(ROOT, "tests", "golden", "traces_3d_refs1")
    names = sorted(os.listdir(gold))
    assert names == ["__ADMMStats_step_0_.txt", "__NewtonIterations_step_0_.txt", "__NewtonStats_step_0_.txt"] and sorted(os.listdir(tmp_path)) == names
    for n in names:
        a, b = gnuplot.read_data(str(tmp_path / n)), gnuplot.read_data(os.path.join(gold, n))
        assert len(a) == len(b) and all(len(x) == len(y) for x, y in zip(a, b)), n
        for x, y in zip(a, b):
            for u, v in zip(x, y):
                assert abs(u - v) <= 1e-9 * max(abs(v), 1e-6), (n, u, v)
    assert open(os.path.join(gold, "__NewtonIterations_step_0_.txt")).read() == open(str(tmp_path / "__NewtonIterations_step_0_.txt")).read()
    ncols = {"__ADMMStats_step_0_.txt": 7, "__NewtonStats_step_0_.txt": 5, "__NewtonIterations_step_0_.txt": 6}
    for n in names:
        assert all(len(r) == ncols[n] for r in gnuplot.read_data(os.path.join(gold, n)))


def test_admm_loop_control_branches():
    """run_admm restates the loop control of 3d_admm.lua:875,1279-1302: a fake convergence doubles `scaling`, resets admm_steps
    and asks for a new J' (callback standing for the Sensitivity re-assembly of 3d:1286-1288); a true convergence breaks."""
    from admm_optim_b200.driver import ObstacleOptim

    class Stub(ObstacleOptim):
        def __init__(self, recs):
            self.P = dict(admmSteps=10, scaling=1.0)
            self.recs, self.calls, self.asked = list(recs), [], []
            self.newton_output, self.trace_dir, self.vNS, self.verbose = False, None, None, False
            self.sensitivity_callback = lambda sc: self.asked.append(sc) or [0.0]
        def begin_step(self):
            self.admm_steps, self.admm_trace, self.p_solver_failure = 0, [], False
        def set_sensitivity(self, j, scaling=None):
            self.calls.append(scaling)
        def admm_iteration(self):
            r = self.recs.pop(0)
            if r is not None:
                self.admm_trace.append((self.admm_steps, r))
                self.admm_steps += 1
            return r

    no = dict(converged=False, fake=False)
    s = Stub([no, dict(converged=True, fake=True), no, dict(converged=True, fake=False), no])
    tr = s.run_admm()
    assert [k for k, _ in tr] == [0, 1, 1, 2]          # after the fake convergence: admm_steps = 0, then the increment at the end of the body
    assert s.P["scaling"] == 2.0 and s.asked == [2.0] and s.calls == [2.0] and len(s.recs) == 1 and not s.p_solver_failure
    s = Stub([no] * 12)
    assert len(s.run_admm()) == 10 and not s.p_solver_failure     # `while admm_steps < admmSteps`: ends without marking the step
    s = Stub([no, None])
    assert len(s.run_admm()) == 1                                   # solver failure: break


# Lua-visible names the reference's hot path needs from the replaced plugins / from the solver glue, with the call sites they were
# taken from (checked against /root/reference when it is present, see the test below)
SHIM_ELEMDISC_CLASSES = ["DeformationEquation", "DeformationEquationRHS", "DeformationEquationLargeProblemRHS",     # 3d:393,407,472
                         "VolumeConstraintSecondDerivative", "SecondDerivativeVolume", "SecondDerivativeBarycenter",     # 3d:559, 2d:564, 3d:577
                         "XBarycenterConstraintSecondDerivative", "MassModel", "LambdaUpdate"]                         # 3d:616,652,677
SHIM_FREE_FUNCTIONS = ["Testing", "ProjectWithSpectralNorm", "MaximumFrobeniusNorm", "MaxSpectralNorm", "VolumeDefect", "BarycenterDefect",
                       "SetZeroAwayFromSubset", "TransformDomainByDisplacement"]                                       # 3d:910,916,1167,1168,817,1333; 2d:901-902
SHIM_GMG_SETTERS = ["set_base_level", "set_base_solver", "set_gathered_base_solver_if_ambiguous", "set_smoother", "set_cycle_type",
                    "set_num_presmooth", "set_num_postsmooth", "set_rap", "set_discretization"]                        # obstacle_optim_3d_util.lua:159-172


def test_ug4_plugin_shim_registers_the_names_the_scripts_call(tmp_path):
    """plugins/ADMMOptimB200/admm_b200_plugin.cpp (the UG4 registry shim over the C ABI) compiles against the stand-in of UG4's
    bridge header, links against libadmm_b200.so, and its InitUGPlugin_ADMMOptimB200 registers every ElemDisc class with every
    setter the drivers call on it, every plugin free function and the GMG / BiCGStab / CG surface of the solver glue."""
    import json
    import re
    import shutil
    if not shutil.which("g++"):
        pytest.skip("no C++ compiler")
    lib_dir = os.path.join(ROOT, "admm_optim_b200")
    if not os.path.exists(os.path.join(lib_dir, "libadmm_b200.so")):
        pytest.skip("libadmm_b200.so not built")
    exe = str(tmp_path / "shim_test")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "ug4_stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "plugins", "ADMMOptimB200", "admm_b200_plugin.cpp"), os.path.join(ROOT, "tests", "ug4_stub", "shim_main.cpp"),
                        "-o", exe, "-L" + lib_dir, "-l:libadmm_b200.so", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    reg = json.loads(out.stdout)
    classes, functions = reg["classes"], set(reg["functions"])
    assert set(SHIM_ELEMDISC_CLASSES) <= set(classes) and all(classes[c]["constructors"] >= 1 for c in SHIM_ELEMDISC_CLASSES)
    assert set(SHIM_FREE_FUNCTIONS) <= functions
    assert set(SHIM_GMG_SETTERS) <= set(classes["B200GeometricMultiGrid"]["methods"])
    assert {"init", "apply", "apply_return_defect", "step", "set_convergence_check"} <= set(classes["B200LinearSolver"]["methods"])
    assert "set_preconditioner" in classes["B200BiCGStab"]["methods"] and "set_preconditioner" in classes["B200CG"]["methods"]
    assert {"set", "has_storage_type_additive", "change_storage_type_to_consistent"} <= set(classes["B200GridFunction"]["methods"])
    assert {"add", "assemble_jacobian", "assemble_defect", "adjust_solution"} <= set(classes["B200DomainDiscretization"]["methods"])
    assert {"B200VecProd", "B200VecScaleAssign", "B200VecScaleAdd2", "B200VecNorm", "B200L2Norm", "B200LoadDomain", "B200CreateRegularHierarchy"} <= functions
    setters = set(classes["B200ElemDisc"]["methods"])
    ref = "/root/reference"
    if os.path.isdir(ref):          # every method the unchanged drivers call on a hot-path ElemDisc object is registered
        called = set()
        pat = re.compile(r"(?:DeformationEquation\w*_ElemDisc|BVolume_ElemDisc|[XYZ]Barycenter_ElemDisc|MassModel_ElemDisc|LambdaUpdate_ElemDisc):(\w+)")
        for f in ("3d_admm.lua", "2d_admm.lua"):
            called |= set(pat.findall(open(os.path.join(ref, f)).read()))
        assert len(called) > 40 and called <= setters, sorted(called - setters)
        src = open(os.path.join(ref, "3d_admm.lua")).read()
        for name in SHIM_ELEMDISC_CLASSES + SHIM_FREE_FUNCTIONS:
            assert re.search(r"\b%s\(" % name, src) or re.search(r"\b%s\(" % name, open(os.path.join(ref, "2d_admm.lua")).read()), name


def test_bench_b200_arm_dry_run_produces_the_full_line():
    """Control flow of bench.py's B200 arm without a GPU (tools/bench_dryrun.py: torch.cuda and the CUDA backend replaced by
    stand-ins on top of the oracle): every leg runs and the JSON line carries every key of the contract and of DESIGN.md section 8.
    Checks the Python of the arm only -- the numbers are meaningless here."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_dryrun.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "clocks", "roofline", "parity", "vcycle_frac_effective", "vcycle_frac_dram", "solve_ms",
                "gmg_init_ms", "assemble_ms", "admm_refs1", "admm_2d_refs2", "elementwise_roofline", "bicgstab_its_per_step"):
        assert key in line, key
    assert line["parity"]["ok"] and "workload" in line["config"] and "error" not in line["elementwise_roofline"]
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.common_config(1)           # identical to the reference arm's for the same --refs
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"]) and line["e2e"]["h2d_bytes_per_step"] > 0


def test_bench_b200_arm_multi_rank_dry_run_gloo_world_size_2():
    """The multi-rank control flow of bench.py (rank bookkeeping, barriers, max-over-ranks timing, the decomposed legs' extra keys, the
    decomposed-vs-undivided parity block on rank 0, the common exit) on 2 gloo ranks with the stand-ins of tools/bench_dryrun.py
    (DRYRUN_DECOMPOSED=1 makes the larger legs claim to be decomposed).  Rank 0 alone prints the line; both ranks exit 0."""
    import json
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29692", WORLD_SIZE="2", DRYRUN_DECOMPOSED="1")
    args = ["--gpus", "2", "--refs", "0", "--roofline-refs", "1", "--admm-refs", "1", "--dim2-refs", "1", "--steps", "1", "--warmup", "1"]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "bench_dryrun.py")] + args, env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=900)
        assert p.returncode == 0, err[-3000:]
        outs.append(out.strip())
    assert outs[1] == ""
    line = json.loads(outs[0].splitlines()[-1])
    assert line["n_gpus"] == 2 and line["parity"]["ok"] and line["parity"]["decomposed"]["ok"]
    assert line["admm_refs1"]["decomposed"] and line["roofline_decomposed"] and "spmv_with_exchange_ms" in line
    assert line["parity"]["decomposed"]["bicgstab_its"] == line["parity"]["decomposed"]["bicgstab_its_undivided"]


def test_vtu_writer_round_trip(tmp_path):
    """admm_optim_b200/vtk.py (VTKOutput of deformation-space functions, 3d_admm.lua:1400-1406): an XML UnstructuredGrid with the
    current coordinates, tetrahedra / triangles and nodal vectors padded to 3 components; read back exactly."""
    from admm_optim_b200 import vtk
    rng = np.random.default_rng(0)
    for dim in (2, 3):
        xyz = rng.standard_normal((7, dim))
        elems = np.array([[0, 1, 2, 3][:dim + 1], [3, 4, 5, 6][:dim + 1]], np.int32)
        u = rng.standard_normal((7, dim))
        path = str(tmp_path / ("u%d.vtu" % dim))
        vtk.write_vtu(path, xyz, elems, {"u": u, "s": u[:, 0]})
        back = vtk.read_vtu(path)
        assert np.array_equal(back["points"][:, :dim], xyz) and np.all(back["points"][:, dim:] == 0)
        assert np.array_equal(back["connectivity"], elems) and list(back["offsets"]) == [dim + 1, 2 * (dim + 1)]
        assert set(back["types"]) == {5 if dim == 2 else 10}
        assert np.array_equal(back["point_data"]["u"][:, :dim], u) and back["point_data"]["u"].shape[1] == 3
        assert np.array_equal(back["point_data"]["s"][:, 0], u[:, 0])


def test_vtkoutput_selection_and_naming_on_the_oracle_backend(tmp_path):
    """vtkWriter:clear_selection(); :select_nodal("u1,u2,u3","u"); :print("u", u, step, step, false) -- 3d_admm.lua:1400-1406."""
    from admm_optim_b200 import vtk
    from admm_optim_b200.driver import ObstacleOptim
    from oracle import ug4_np
    ug = ug4_np.Backend()
    p = ObstacleOptim(ug, 3, numRefs=0, grid=GRID3D).setup()
    p.u.from_numpy(np.random.default_rng(2).standard_normal(p.u.v.size))
    w = vtk.VTKOutput(ug)
    w.clear_selection()
    w.select_nodal(p.ucmps, "u")
    path = w.print(str(tmp_path / "u"), p.u, 3, 3, False)
    assert os.path.basename(path) == "u_t0003.vtu"
    back = vtk.read_vtu(path)
    assert np.array_equal(back["points"], p.dom.top.xyz) and np.array_equal(back["connectivity"], p.dom.top.elems)
    assert np.array_equal(back["point_data"]["u"], p.u.to_numpy().reshape(-1, 3))
    w.clear_selection()
    w.select_nodal("u2", "uy")
    back = vtk.read_vtu(w.print(str(tmp_path / "uy"), p.u))
    assert np.array_equal(back["point_data"]["uy"][:, 0], p.u.to_numpy().reshape(-1, 3)[:, 1])


def _special_set(g):
    e = {(int(a), int(b), int(s)) for (a, b), s in zip(np.sort(g["sp_edges"], axis=1), g["sp_edges_sub"])}
    f = {(int(a), int(b), int(c), int(s)) for (a, b, c), s in zip(np.sort(g["sp_faces"], axis=1), g["sp_faces_sub"])} if len(g["sp_faces"]) else set()
    return e, f


@pytest.mark.parametrize("grid,level", [(GRID3D, 0), (GRID3D, 1), (GRID2D, 2)])
def test_save_grid_level_to_file_round_trip(grid, level, tmp_path):
    """SaveGridLevelToFile (3d_admm.lua:795): the written .ugx read back by LoadDomain reproduces the level bit for bit
    (coordinates, elements, subsets of vertices / boundary edges / boundary faces), by the oracle's reader too, and refining
    the re-read grid gives the next level of the original hierarchy."""
    from admm_optim_b200 import ug4
    from oracle import mesh_np as M
    dom = _host_domain(grid, level + 1)
    ug = dom.ug
    name = ug.SaveGridLevelToFile(dom.grid(), dom.subset_handler(), level, str(tmp_path / ("Mesh_lev%d_step1.ugx" % level)))
    g = dom.get_grid_dict(level)
    back = ug4.Domain(ug)
    ug.LoadDomain(back, name)
    h = back.get_grid_dict(0)
    assert h["dim"] == g["dim"] and h["subset_names"] == g["subset_names"]
    for k in ("xyz", "elems", "vsub", "esub"):
        assert np.array_equal(h[k], g[k]), k
    assert _special_set(h) == _special_set(g)
    o = M.load_ugx(name)
    assert np.array_equal(o.xyz, g["xyz"]) and np.array_equal(o.elems, g["elems"]) and np.array_equal(o.vsub, g["vsub"])
    ug.util.refinement.CreateRegularHierarchy(back, 1, False, None)
    a, b = back.get_level(1), dom.get_level(level + 1)
    assert np.array_equal(a["xyz"], b["xyz"]) and np.array_equal(a["elems"], b["elems"]) and np.array_equal(a["vsub"], b["vsub"])
    # every edge / face / volume sits in exactly one subset, as in the shipped grids
    txt = open(name).read()
    n_edges = len(re.search(r"<edges>(.*?)</edges>", txt, re.S).group(1).split()) // 2
    listed = sum(len(m.split()) for m in re.findall(r"<subset name.*?</subset>", txt, re.S) for m in re.findall(r"<edges>(.*?)</edges>", m, re.S))
    assert listed == n_edges == dom.level_info(level)["nedges"]


def _build_shim(tmp_path):
    import shutil
    lib_dir = os.path.join(ROOT, "admm_optim_b200")
    if not shutil.which("g++") or not os.path.exists(os.path.join(lib_dir, "libadmm_b200.so")):
        pytest.skip("no C++ compiler / library not built")
    exe = str(tmp_path / "shim_test")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "ug4_stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "plugins", "ADMMOptimB200", "admm_b200_plugin.cpp"), os.path.join(ROOT, "tests", "ug4_stub", "shim_main.cpp"),
                        "-o", exe, "-L" + lib_dir, "-l:libadmm_b200.so", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.parametrize("grid,refs", [(GRID3D, 1), (GRID2D, 2)])
def test_ug4_plugin_shim_writers(grid, refs, tmp_path):
    """The output functions of the shim (B200SaveGridLevelToFile, the .vtu writer behind B200VTKOutput) run without a GPU on a
    host-only domain / synthetic data: the .ugx of a refined level equals the Python mirror's file entity by entity after
    re-reading both, and the .vtu parses with the reader of admm_optim_b200/vtk.py."""
    from admm_optim_b200 import ug4
    from admm_optim_b200.vtk import read_vtu
    exe = _build_shim(tmp_path)
    dom = _host_domain(grid, refs)
    ug = dom.ug
    base = ug.SaveGridLevelToFile(dom.grid(), dom.subset_handler(), 0, str(tmp_path / "level0.ugx"))
    py = ug.SaveGridLevelToFile(dom.grid(), dom.subset_handler(), refs, str(tmp_path / "py.ugx"))
    cc = str(tmp_path / "cc.ugx")
    r = subprocess.run([exe, "ugx", base, str(refs), str(refs), cc], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = ug4.Domain(ug), ug4.Domain(ug)
    ug.LoadDomain(a, py)
    ug.LoadDomain(b, cc)
    ga, gb, g = a.get_grid_dict(0), b.get_grid_dict(0), dom.get_grid_dict(refs)
    for k in ("xyz", "elems", "vsub", "esub", "sp_edges", "sp_edges_sub", "sp_faces", "sp_faces_sub"):
        assert np.array_equal(ga[k], gb[k]), k
    assert np.array_equal(gb["xyz"], g["xyz"]) and np.array_equal(gb["elems"], g["elems"]) and gb["subset_names"] == g["subset_names"]
    assert _special_set(gb) == _special_set(g)
    # the two files list the same edges / faces in the same order
    import re as _re
    for tag in ("edges", "triangles"):
        ta = _re.search(r"<%s>(.*?)</%s>" % (tag, tag), open(py).read(), _re.S).group(1).split()
        tb = _re.search(r"<%s>(.*?)</%s>" % (tag, tag), open(cc).read(), _re.S).group(1).split()
        assert ta == tb, tag
    out = str(tmp_path / "f.vtu")
    assert subprocess.run([exe, "vtu", out]).returncode == 0
    v = read_vtu(out)
    assert v["points"].shape == (5, 3) and v["connectivity"].tolist() == [[0, 1, 2, 3], [1, 2, 3, 4]] and set(v["types"]) == {10}
    vals = (0.1 * np.arange(15) - 0.3).reshape(5, 3)
    assert np.array_equal(v["point_data"]["u"], vals) and np.array_equal(v["point_data"]["first"][:, 0], vals[:, 0])


@pytest.mark.parametrize("grid,refs", [(GRID3D, 2), (GRID2D, 4)])
def test_native_pattern_and_incidence_match_numpy(grid, refs):
    """The (multi-threaded) host builders of the P1 block pattern and of the vertex -> element incidence give exactly the NumPy
    statement: rows = sorted unique vertex pairs sharing an element; incidence = elements ascending per vertex."""
    import ctypes as C
    from admm_optim_b200 import partition as P, ug4
    dom = _host_domain(grid, refs)
    for level in range(refs + 1):
        lv = dom.get_level(level)
        nv, el = len(lv["xyz"]), lv["elems"]
        rp, ci = dom.level_pattern(level)
        assert np.array_equal(P.csr_keys(rp, ci, nv), P.pattern_keys(el, nv))
        ptr, idx = np.empty(nv + 1, np.int32), np.empty(el.size, np.int32)
        ug4.call("ab_domain_level_incidence", dom.h, level, ptr.ctypes.data_as(C.POINTER(C.c_int32)), idx.ctypes.data_as(C.POINTER(C.c_int32)))
        order = np.argsort(el.ravel(), kind="stable")                   # by vertex, then by position in the element list = element ascending
        assert np.array_equal(idx, (order // el.shape[1]).astype(np.int32))
        assert np.array_equal(np.diff(ptr), np.bincount(el.ravel(), minlength=nv))


@pytest.mark.parametrize("grid,refs", [(GRID3D, 3), (GRID2D, 4)])
def test_edges_from_parent_equal_generic_edges(grid, refs, monkeypatch):
    """Refined levels take their edge list from the parent level (halves of coarse edges + midpoint-midpoint edges of the inner
    children) instead of from all element edges: same hierarchy (vertex numbering of the next level = edge order) and same
    patterns as the generic path."""
    a = _host_domain(grid, refs)
    monkeypatch.setenv("ADMM_B200_GENERIC_EDGES", "1")
    b = _host_domain(grid, refs)
    for level in range(refs + 1):
        la, lb = a.get_level(level), b.get_level(level)
        for k in ("xyz", "elems", "vsub", "parent_a", "parent_b"):
            assert np.array_equal(la[k], lb[k]), (level, k)
        assert a.level_info(level)["nedges"] == b.level_info(level)["nedges"]
        for x, y in zip(a.level_pattern(level), b.level_pattern(level)):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("mode", ["raise", "hang"])
def test_bench_line_survives_a_failing_or_hanging_secondary_leg(mode):
    """The headline of bench.py's B200 arm is never lost to a secondary leg: an exception in a leg, or a leg that exceeds
    --legs-timeout (a collective waiting for a lost peer), still ends with exit code 0 and the JSON line carrying the headline,
    the legs measured before, and `secondary_legs_error` naming the leg."""
    import json
    env = dict(os.environ, DRYRUN_FAIL_IN_2D_LEG=mode)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_dryrun.py"), "--refs", "1", "--roofline-refs", "1", "--admm-refs", "0", "--dim2-refs", "2",
                        "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--legs-timeout", "10" if mode == "hang" else "600"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["parity"]["ok"]
    assert "roofline" in line and "vcycle_ms" in line                      # the leg that ran before the failing one is kept
    assert "admm_2d_refs2" not in line and "admm_2d_refs2" in line["secondary_legs_error"]
    assert ("injected failure" if mode == "raise" else "legs-timeout") in line["secondary_legs_error"]


def test_interface_exchange_protocol_on_the_host(tmp_path):
    """The source of the NVLink peer-memory interface exchange (csrc/iface_xchg.cuh: slot tables + k_iface_xchg) compiled for the
    CPU -- one std::thread per CUDA thread, C++ atomics for the flag traffic (tests/cuda_host_shim/xchg_emulation.cpp): 4-5 ranks
    drifting up to one exchange apart over many launches, capped grids with grid-stride phases, vertices on up to 3 ranks.
    Consistent copies must be bitwise equal to the rank-ordered sum; under ThreadSanitizer no window access may race; and a mutant
    with single-buffered windows must be caught (the harness can fail)."""
    import shutil
    if not shutil.which("g++"):
        pytest.skip("no C++ compiler")
    src = os.path.join(ROOT, "tests", "cuda_host_shim", "xchg_emulation.cpp")
    inc = os.path.join(ROOT, "admm_optim_b200", "csrc")
    base = ["g++", "-std=c++20", "-O1", "-g", "-pthread", "-ffp-contract=off", "-w"]

    def build(out, include, tsan):
        r = subprocess.run(base + (["-fsanitize=thread"] if tsan else []) + ["-I" + include, src, "-o", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        return out

    plain = build(str(tmp_path / "xchg"), inc, False)
    for cfg in (["4", "3", "60", "8", "3", "8"], ["5", "2", "300", "12", "2", "4"], ["2", "3", "40", "6", "1", "4"], ["3", "1", "50", "9", "4", "2"]):
        r = subprocess.run([plain] + cfg, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout[-2000:]
    probe = subprocess.run(base + ["-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "probe")], input="int main(){return 0;}", capture_output=True, text=True)
    if probe.returncode != 0:
        return                                              # no ThreadSanitizer runtime on this machine: the functional part stands
    tsan = build(str(tmp_path / "xchg_tsan"), inc, True)
    r = subprocess.run([tsan, "4", "3", "60", "12", "3", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "data race" not in r.stderr and r.stdout.startswith("OK"), (r.stdout + r.stderr)[-3000:]
    mut = tmp_path / "mut"
    mut.mkdir()
    text = open(os.path.join(inc, "iface_xchg.cuh")).read()
    assert "const int parity = (int)(epoch & 1ull);" in text
    (mut / "iface_xchg.cuh").write_text(text.replace("const int parity = (int)(epoch & 1ull);", "const int parity = 0;"))
    bad = build(str(tmp_path / "xchg_mut"), str(mut), True)
    r = subprocess.run([bad, "4", "3", "60", "30", "3", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 or "data race" in r.stderr, "single-buffered windows went unnoticed"
