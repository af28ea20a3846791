"""Host layer checks that need no device: the stand-alone driver's object registry and its
hit-subset input reader (host/shim/hit.C), on this repository's inputs and - in the build
container only - on every input file the reference ships."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "marlin_b200", "marlin_b200-opt")
INP = os.path.join(ROOT, "tests", "inputs")
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.exists(APP), reason="host driver not built (make)")


def run(*args, ok=True):
    r = subprocess.run([APP, *args], capture_output=True, text=True, timeout=60)
    assert (r.returncode == 0) == ok, r.stdout + r.stderr
    return r


def tree(*args):
    out = {}
    for line in run(*args, "--parse-only").stdout.splitlines():
        if " = " in line:
            k, v = line.split(" = ", 1)
            out[k] = v
    return out


def test_registry_holds_the_hot_path_objects():
    """registerMooseObject names of SURVEY.md 8(a)/(b) (src/base/MarlinApp.C:94-172 wiring)."""
    names = set(run("--list-objects").stdout.split())
    for n in ["TensorProblem", "ComputeGroup", "ForwardFFT", "InverseFFT", "ParsedCompute",
              "ReciprocalLaplacianFactor", "ReciprocalLaplacianSquareFactor", "RandomTensor",
              "ConstantTensor", "ConstantReciprocalTensor", "FFTGradient", "FFTGradientSquare",
              "FFTSemiImplicit", "AdamsBashforthMoulton", "SemiImplicitSolver", "ForwardEulerSolver",
              "ETDRK4Solver", "FFTMechanics", "HyperElasticIsotropic", "RankTwoIdentity",
              "TensorAveragePostprocessor", "TensorIntegralPostprocessor",
              "TensorExtremeValuePostprocessor", "TensorIntegralChangePostprocessor",
              "SemiImplicitCriticalTimeStep"]:
        assert n in names, n


def test_fparse_units_and_top_level_variables():
    t = tree("-i", f"{INP}/etdrk4_decay.i")
    assert float(t["Domain/xmax"]) == pytest.approx(6.283185307179586, abs=0)
    assert t["dt"] == "10" and t["Executioner/dt"] == "10"          # ${units 10 s}; dt = ${dt}
    assert t["TensorComputes/Solve/u_exact/expression"] == "u0*exp(-0.05*1.0^2*t)"
    assert t["TensorSolver/substeps"] == "1"


def test_command_line_overrides():
    t = tree("-i", f"{INP}/abm_diagonal.i", "ss=20", "cs=2", "order=4", "Domain/nx=64",
             "Executioner/num_steps=3")
    assert t["TensorSolver/substeps"] == "20" and t["TensorSolver/corrector_steps"] == "2"
    assert t["TensorSolver/predictor_order"] == "4" and t["Domain/nx"] == "64"
    assert t["Outputs/file_base"] == "abm_diagonal_20_2_4" and t["Executioner/num_steps"] == "3"
    r = run("-i", f"{INP}/abm_diagonal.i", "--parse-only", ok=False)  # ${ss} undefined
    assert "ss" in r.stderr


def test_active_filter_and_quoted_lists():
    out = run("-i", f"{INP}/ch2d_gold.i", "--parse-only").stdout
    assert "[AuxKernels]" in out and "[AuxKernels/c]" not in out      # active = ''
    t = tree("-i", f"{INP}/ch2d_gold.i")
    assert t["TensorComputes/Solve/cahn_hilliard/Mbarmubar/inputs"] == "Mbar mubar"


def test_syntax_errors_are_reported_with_line(tmp_path):
    p = tmp_path / "bad.i"
    p.write_text("[Domain]\n  dim = 2\n[TensorComputes]\n")
    r = run("-i", str(p), "--parse-only", ok=False)
    assert "bad.i" in r.stderr
    p.write_text("[Domain]\n  dim = ${nope}\n[]\n")
    r = run("-i", str(p), "--parse-only", ok=False)
    assert "nope" in r.stderr and ":2" in r.stderr


def test_include(tmp_path):
    (tmp_path / "a.i").write_text("n = 7\n!include b.i\n")
    (tmp_path / "b.i").write_text("[Domain]\n  nx = ${n}\n[]\n")
    assert tree("-i", str(tmp_path / "a.i"))["Domain/nx"] == "7"


def test_no_device_is_a_loud_error():
    """There is no CPU path: building the objects without a GPU must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    r = run("-i", f"{INP}/ch2d_gold.i", ok=False)
    assert "no CPU path" in r.stderr


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_every_reference_input_parses():
    files = [f for d in ("examples", "benchmarks", "test/tests")
             for f in glob.glob(f"{REF}/{d}/**/*.i", recursive=True)]
    assert len(files) > 50
    for f in files:
        r = subprocess.run([APP, "-i", f, "--parse-only", "ss=10", "cs=0", "order=2"],
                           capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, f + "\n" + r.stderr


# tests/inputs/ref/<file> -> path below the reference root
VERBATIM = {
    "cahnhilliard.i": "test/tests/cahnhilliard", "cahnhilliard_explicit.i": "test/tests/cahnhilliard",
    "diagonal.i": "test/tests/solvers", "coupled.i": "test/tests/solvers", "nl_coupled.i": "test/tests/solvers",
    "etdrk4_diffusion.i": "test/tests/solvers", "mech3d.i": "test/tests/mechanics", "mech.i": "test/tests/mechanics",
    "1a_solver.i": "benchmarks/01_spinodal_decomposition", "2a.i": "benchmarks/02_oswald_ripening",
    "cahnhilliard2.i": "examples/cahn_hilliard", "gradient.i": "test/tests/gradient",
    "rotating_grain_secant.i": "test/tests/tensor_compute", "KKS_no_flux_bc.i": "test/tests/kks",
    "postprocessors.i": "test/tests/postprocessors",
}


def test_vendored_reference_inputs_parse_and_are_listed():
    """tests/inputs/ref/ holds exactly the files of VERBATIM, and the reader resolves each of them."""
    have = sorted(os.path.basename(f) for f in glob.glob(f"{INP}/ref/*.i"))
    assert have == sorted(VERBATIM)
    for f in have:
        run("-i", f"{INP}/ref/{f}", "--parse-only", "ss=10", "cs=0", "order=2")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_vendored_reference_inputs_are_verbatim():
    """Byte for byte the reference's files (tests/test_gpu_ref_inputs.py runs them on the GPU)."""
    for f, d in VERBATIM.items():
        assert open(f"{INP}/ref/{f}", "rb").read() == open(f"{REF}/{d}/{f}", "rb").read(), f


def test_unused_parameters_error_unless_allowed():
    """MOOSE's default ERROR_UNUSED (moose/framework/src/base/MooseApp.C:477, Builder.C:361-399): the stale
    `history_size` / `spectral_solve_substeps` of benchmarks/01_spinodal_decomposition/1a_solver.i are rejected when the
    objects are built, with the reference's hint; --allow-unused / -w turn them into warnings.  Needs no device up to
    the error: parameters are filled before the context is created - checked on the GPU box in test_gpu_ref_inputs.py."""
    r = run("-i", f"{INP}/ref/1a_solver.i", "--parse-only", "-w")
    assert r.returncode == 0


def test_xdmf_writer_selftest(tmp_path):
    """XDMFTensorOutput's writer (src/tensor_outputs/XDMFTensorOutput.C:118-221 skeleton, :266-355 data, :358-426
    per-frame XML, :529-553 periodic continuation for NODE, buildAttributeNames :654-668) on synthetic host data:
    document structure as in the reference's gold cahnhilliard.xmf (binary DataItems instead of HDF), data files
    with the extension / transpose applied."""
    import xml.etree.ElementTree as ET

    import numpy as np
    r = subprocess.run([APP, "--xdmf-selftest", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(tmp_path / "selftest.xmf").read()
    # the skeleton, character for character in the layout pugixml gives the reference (tabs, attribute order)
    assert text.startswith('<?xml version="1.0"?>\n<Xdmf xmlns:xi="http://www.w3.org/2003/XInclude" Version="2.2">\n\t<Domain>\n'
                           '\t\t<Topology TopologyType="2DCoRectMesh" Dimensions="4 3" />\n\t\t<Geometry Type="ORIGIN_DXDY">\n'
                           '\t\t\t<DataItem Format="XML" Dimensions="2">0 -1</DataItem>\n'
                           '\t\t\t<DataItem Format="XML" Dimensions="2">0.5 0.25</DataItem>\n\t\t</Geometry>\n'
                           '\t\t<Grid Name="TimeSeries" GridType="Collection" CollectionType="Temporal">\n'
                           '\t\t\t<Grid Name="T0" GridType="Uniform">\n\t\t\t\t<Time Value="0" />\n'
                           '\t\t\t\t<xi:include xpointer="xpointer(//Xdmf/Domain/Topology)" />\n')
    assert '<Time Value="0.0030000000000000001" />' in text and text.endswith("\t\t</Grid>\n\t</Domain>\n</Xdmf>\n")
    root = ET.fromstring(text)
    grids = root.find("Domain").find("Grid").findall("Grid")
    assert [g.get("Name") for g in grids] == ["T0", "T1"]
    attrs = [(a.get("Name"), a.get("Center"), a.find("DataItem").get("Dimensions"), a.find("DataItem").get("Format"))
             for a in grids[1].findall("Attribute")]
    assert attrs == [("c", "Node", "4 3", "Binary"), ("disp_x", "Node", "4 3", "Binary"), ("disp_y", "Node", "4 3", "Binary"),
                     ("mu", "Cell", "3 2", "Binary")]
    rd = lambda name, shape: np.fromfile(tmp_path / name, dtype="<f8").reshape(shape)
    c = np.arange(10.0, 16.0).reshape(3, 2)
    ext = np.concatenate([np.concatenate([c, c[:1]], 0), np.concatenate([c, c[:1]], 0)[:, :1]], 1)   # periodic continuation
    assert np.array_equal(rd("selftest.c.0.bin", (4, 3)), ext)
    assert np.array_equal(rd("selftest_t.c.0.bin", (3, 4)), ext.T)                                  # transpose = true swaps x and y
    assert np.array_equal(rd("selftest.c.1.bin", (4, 3)), ext + 100)
    assert np.array_equal(rd("selftest.mu.0.bin", (3, 2)), np.arange(20.0, 26.0).reshape(3, 2))
    assert np.array_equal(rd("selftest.disp_y.1.bin", (4, 3)), np.arange(112.0, 124.0).reshape(4, 3))  # oversized nodal: as is
    tt = open(tmp_path / "selftest_t.xmf").read()
    assert 'Dimensions="3 4" />' in tt and ">-1 0<" in tt and ">0.25 0.5<" in tt                     # mapped axes


def test_every_repository_input_parses_and_names_registered_types():
    """tests/inputs/*.i (the restated reference inputs the GPU host tests run): the hit reader accepts them
    and every `type =` of a tensor object is registered with the factory (--list-objects)."""
    import re
    registered = set(run("--list-objects").stdout.split())
    files = sorted(glob.glob(f"{INP}/*.i"))
    assert len(files) >= 15
    for f in files:
        r = subprocess.run([APP, "-i", f, "--parse-only", "ss=10", "cs=0", "order=2", "smooth=SHARP"], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, f + "\n" + r.stderr
        block = None
        for line in r.stdout.splitlines():
            m = re.match(r"(\S+)/type = (\S+)$", line)
            if not m:
                continue
            path, typ = m.groups()
            top = path.split("/")[0]
            if top in ("TensorComputes", "TensorSolver", "TensorOutputs") or (top == "Postprocessors" and typ.startswith(("Tensor", "Reciprocal", "SemiImplicit", "ComputeGroup"))):
                assert typ in registered, f"{f}: {path} uses unregistered type {typ}"


def test_smooth_rectangle_kernel_expression_matches_gold():
    """SmoothRectangleCompute (src/tensor_computes/SmoothRectangleCompute.C:60-131) is ONE generated kernel on the
    device; its expression, as the host object builds it, must (a) compile with NVRTC for sm_100a and (b) give the
    values of the reference's gold smooth_rectangle.h5 when the C++ expression evaluator (the AST the code generator
    walks) is run at the cell centres of that case."""
    import numpy as np

    from marlin_b200 import capi
    g = np.load(os.path.join(ROOT, "tests", "golden", "smooth_rectangle_h5.npz"))
    consts = dict(x1=5.0, x2=15.0, y1=5.0, y2=15.0, z1=0.0, z2=0.0, w=1.0, w2=0.5, vin=-1.0, vout=3.0)
    centre = np.linspace(0.1, 19.9, 100)                      # linspace(min + dx/2, max - dx/2, n), DomainAction.C:227-338
    pts = [(i, j) for i in (0, 22, 23, 24, 25, 26, 27, 50, 73, 74, 75, 76, 77, 99) for j in (3, 23, 25, 26, 50, 74, 75, 77)]
    for name, w, profile in [("sharp", 0.0, "NONE"), ("cos", 1.0, "COS"), ("tanh", 1.0, "TANH")]:
        expr = run("--smooth-rectangle-expr", "2", str(w), profile).stdout.strip()
        src = capi.expr_check(expr, constants={**consts, "w": w, "w2": w / 2, "pi": np.pi}, extra_symbols=True, expand=capi.EXPAND_REAL)
        assert "mrl_expr_u32" in src
        for i, j in pts:
            v = capi.expr_constant(expr, {**consts, "w": w, "w2": w / 2, "x": centre[i], "y": centre[j]})
            assert abs(v - g[name][i, j]) < 1e-13, (name, i, j, v, g[name][i, j])
    e3 = run("--smooth-rectangle-expr", "3", "0.5", "TANH").stdout.strip()
    assert "dist_z" in e3 and capi.expr_check(e3, constants={**consts, "pi": np.pi}, extra_symbols=True, expand=capi.EXPAND_REAL)
    assert run("--smooth-rectangle-expr", "2", "1", "NONE").stdout.strip() == "vout"


# object types the reference's inputs use that this drop-in leaves out ("out of scope"), with why
OUT_OF_SCOPE_TYPES = {
    "FiniteDifferenceLaplacian": "real-space finite differences (test/tests/real_space), not the spectral path",
    "RealSpaceForwardEuler": "real-space finite differences",
    "PlainTensorBuffer": "real-space ghost-layer buffers",
    "GradientTensor": "needs NEML2 (mooseError without it, src/tensor_computes/GradientTensor.C:42-44)",
    "NEML2TensorCompute": "needs NEML2 (un-vendored)",
    "LibtorchGibbsEnergy": "TorchScript model file",
    "ParsedTensor": "not registered in the reference either (stale inputs test.i / sineic.i)",
}


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_every_spectral_object_type_of_the_reference_inputs_is_registered():
    """Every `type =` under [TensorComputes] / [TensorSolver] / [TensorOutputs] / [TensorBuffers] in the inputs the
    reference ships (examples, benchmarks, test/tests) is registered with this driver's factory, except the lattice
    Boltzmann objects (LBM*) and the short list above."""
    import re
    registered = set(run("--list-objects").stdout.split())
    missing = {}
    for d in ("examples", "benchmarks", "test/tests"):
        for f in glob.glob(f"{REF}/{d}/**/*.i", recursive=True):
            r = subprocess.run([APP, "-i", f, "--parse-only", "ss=10", "cs=0", "order=2"], capture_output=True, text=True, timeout=60)
            for line in r.stdout.splitlines():
                m = re.match(r"(\S+)/type = (\S+)$", line)
                if m and m.group(1).split("/")[0] in ("TensorComputes", "TensorSolver", "TensorOutputs", "TensorBuffers"):
                    typ = m.group(2)
                    if typ not in registered and not typ.startswith("LBM") and typ not in OUT_OF_SCOPE_TYPES:
                        missing.setdefault(typ, f)
    assert not missing, missing


def test_hdf5_writer_against_the_reader_that_reads_libhdf5_files(tmp_path):
    """XDMFTensorOutput with enable_hdf5 = true stores every field as a one-chunk deflate-9 dataset like the reference's
    addDataToHDF5 (src/tensor_outputs/XDMFTensorOutput.C:572-651).  The image has no libhdf5: host/shim/h5lite.C writes the
    file structures itself and tests/h5lite.py reads them back - the same reader that parses the gold files libhdf5 wrote
    (next test) - with the structure versions of those gold files."""
    import numpy as np

    import h5lite
    r = subprocess.run([APP, "--xdmf-selftest", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    h = h5lite.H5File(tmp_path / "selftest_h5.h5")
    assert h.keys() == ["c.0", "c.1", "disp_x.0", "disp_x.1", "disp_y.0", "disp_y.1", "mu.0", "mu.1"]
    assert (h.leaf_k, h.internal_k) == (4, 16)
    for name in h.keys():
        want = np.fromfile(tmp_path / f"selftest_t.{name}.bin", dtype="<f8")      # the binary writer on the same fields
        d = h.info(name)
        assert d["dtype"] == "<f8" and d["layout_class"] == 2 and d["filters"] == [(1, (9,))] and d["chunk"][:-1] == d["shape"]
        assert np.array_equal(h.read(name).ravel(), want)
    assert h.info("c.0")["shape"] == (3, 4) and h.info("mu.0")["shape"] == (2, 3)
    text = open(tmp_path / "selftest_h5.xmf").read()
    assert f'<DataItem DataType="Float" Dimensions="2 3" Format="HDF">{tmp_path}/selftest_h5.h5:/mu.1</DataItem>' in text
    # 300 float32 datasets: symbol table nodes of 8 entries under a two-level group B-tree, flushed while growing
    m = h5lite.H5File(tmp_path / "selftest_many.h5")
    assert len(m.keys()) == 300
    for k in (0, 7, 150, 299):
        a = m.read(f"field_{k}.0")
        assert a.dtype == np.float32 and a.shape == (2, 3, 4)
        assert np.array_equal(a.ravel(), k + 0.5 * np.arange(24, dtype=np.float32))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_h5lite_reader_parses_the_reference_gold_files():
    """tests/h5lite.py on the files libhdf5 wrote for the reference's gold results: the datasets equal the fixtures extracted
    earlier (tests/golden/make_golden.py), and their structure is what host/shim/h5lite.C reproduces."""
    import numpy as np

    import h5lite
    g = np.load(f"{ROOT}/tests/golden/ch2d_slab_rank1_h5.npz")["c"]
    h = h5lite.H5File(f"{REF}/test/tests/cahnhilliard/gold/cahnhilliard.rank0001.h5")
    assert h.sb_version == 0 and (h.leaf_k, h.internal_k) == (4, 16)
    for i in range(11):
        assert np.array_equal(h.read(f"c.{i}"), g[i])
    d = h.info("c.0")
    assert (d["dataspace_version"], d["dtype_version"], d["fill_version"], d["filter_version"], d["layout_version"]) == (1, 1, 2, 1, 3)
    assert d["filters"] == [(1, (9,))] and d["chunk"] == (20, 10, 8)
    m = h5lite.H5File(f"{REF}/test/tests/mechanics/gold/mech3d.h5")
    gm = np.load(f"{ROOT}/tests/golden/mech3d_h5.npz")
    assert np.array_equal(m.read("sV.2"), gm["sV"][2].transpose(2, 1, 0))   # stored with transpose = true (x <-> z)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_process_group_rendezvous(world):
    """host/shim/comm: the MPI-free process group of the stand-alone driver ([Domain] parallel_mode = FFT_SLAB / FFT_PENCIL
    run one process per GPU).  Ranks meet over TCP through the torchrun-style environment; allgather keeps rank order,
    allreduce sums / minimises / maximises in rank order on every rank (so that all ranks take the same decisions), a rank
    that never shows up is an error, not a hang."""
    port = 29400 + (os.getpid() * 3 + world) % 500
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MRL_COMM_PORT=str(port))
        procs.append(subprocess.Popen([APP, "--comm-selftest"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=60) for p in procs]
    gathered = " ".join(f"{r} {10 * r} 7" for r in range(world))
    tri = world * (world + 1) // 2
    for r, (p, (so, se)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, se
        assert f"rank {r} of {world} local {r} gathered {gathered} sum {tri} min 1 max {world} big" in so, so
        assert abs(float(so.split()[-1]) - 0.001 * tri) < 1e-15


def test_process_group_missing_rank_is_an_error():
    """Rank 0 of a world of two whose peer never starts gives up with a message instead of waiting forever."""
    port = 29350 + os.getpid() % 40
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MRL_COMM_PORT=str(port), MRL_COMM_TIMEOUT="1")
    r = subprocess.run([APP, "--comm-selftest"], env=env, capture_output=True, text=True, timeout=30)
    assert r.returncode != 0 and "only 1 of 2 ranks reached the rendezvous" in r.stderr
