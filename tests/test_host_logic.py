"""CPU-side checks: the C-ABI library loads and exports everything include/sttode.h declares, the Python
mirror has the reference's state_dict layout and option behaviour, and nothing silently falls back."""
import os
import re

import pytest
import torch

import sttode_b200._lib as L
from oracle import sttode_oracle as O
from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sttode.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sttode_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sttode.h but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in sttode_b200/_lib.py"
    assert set(L.SIGNATURES) == set(names)
    assert lib.sttode_version() == 100


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(L.MhgsaWeights) == 4 * 8
    assert C.sizeof(L.OdeWeights) == 17 * 8                      # 16 tensors + the optional packed tensor-core images
    assert C.sizeof(L.EncoderWeights) == (9 + 17 + 1) * 8        # + the optional folded front-end table
    assert C.sizeof(L.DecomposeWeights) == (6 + 12) * 8
    assert C.sizeof(L.ModelWeights) == 8 + (27 + 36 + 2) * 8
    # gradient structs mirror the weight structs field for field (include/sttode.h)
    assert C.sizeof(L.MhgsaGrads) == C.sizeof(L.MhgsaWeights) and C.sizeof(L.OdeGrads) == C.sizeof(L.OdeWeights) - 8
    assert C.sizeof(L.DecomposeGrads) == C.sizeof(L.DecomposeWeights) and C.sizeof(L.Mlp3Grads) == C.sizeof(L.Mlp3Weights)
    assert [f[0] for f in L.OdeGrads._fields_] == [f[0] for f in L.OdeWeights._fields_ if f[0] != "packed"]
    assert [f[0] for f in L.DecomposeGrads._fields_] == [f[0] for f in L.DecomposeWeights._fields_]


def make_args(dataset="eth", tp=8, tf=12):
    return ref_shim.make_args(dataset, tp, tf)


@pytest.mark.parametrize("tp,tf", [(8, 12), (5, 10), (10, 20)])
def test_state_dict_layout_matches_reference_spec(tp, tf):
    from sttode_b200.net import STTODENet
    m = STTODENet(make_args("nba", tp, tf), torch.device("cpu"))
    sd = m.state_dict()
    spec = O.state_dict_spec(tp, tf)
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    m.load_state_dict(O.make_state_dict(3, tp, tf), strict=True)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_seeded_init_equals_reference():
    """Same constructor order and init calls => torch.manual_seed(s) gives the reference's initial weights."""
    from sttode_b200.net import STTODENet
    ref = ref_shim.install()
    torch.manual_seed(1)
    a = ref.sttode.STTODENet(make_args(), torch.device("cpu")).state_dict()
    torch.manual_seed(1)
    b = STTODENet(make_args(), torch.device("cpu")).state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_no_cpu_fallback():
    from sttode_b200.mhgsa import Hyp_mhsa
    from sttode_b200 import pmath
    m = Hyp_mhsa(64, 8).eval()
    x = torch.randn(3, 2, 64)
    with pytest.raises(L.SttodeError):
        m(x, x, x)
    with pytest.raises(L.SttodeError):
        pmath.project(torch.randn(4, 8))


def test_unsupported_options_raise():
    from sttode_b200.mhgsa import Hyp_mhsa, hyp_mhsa
    with pytest.raises(NotImplementedError):
        Hyp_mhsa(64, 8, add_bias_kv=True)
    with pytest.raises(NotImplementedError):
        Hyp_mhsa(64, 8, kdim=32)
    m = Hyp_mhsa(64, 8)
    x = torch.randn(3, 2, 64)
    with pytest.raises(NotImplementedError):
        m(x, x, x, attn_mask=torch.zeros(3, 3))
    with pytest.raises(NotImplementedError):
        m(x, x, x, key_padding_mask=torch.zeros(2, 3, dtype=torch.bool))
    from sttode_b200.net import STTODENet
    a = make_args()
    a.learn_prior = True
    with pytest.raises(NotImplementedError):
        STTODENet(a, torch.device("cpu"))
    with pytest.raises(NotImplementedError):
        STTODENet(make_args(), torch.device("cpu")).train().forward()


def test_compat_registers_reference_module_names():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import sttode_b200.compat as c; c.install();"
            "from model.STTODE import STTODENet; from hyptransformerlib import Hyp_mhsa; from ode_demo import ODEG_Encoder;"
            "from core.manifolds import Oblique; import hyptorch.pmath as p; print(STTODENet.__module__, p.__name__)") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "sttode_b200.net" in out.stdout


def test_synth_is_deterministic():
    from sttode_b200 import synth
    a, _ = synth.eth_scene(3, 7)
    b, _ = synth.eth_scene(3, 7)
    assert torch.equal(a, b) and a.shape == (7, 2, 8)
    d = synth.nba_batch(1, 4)
    assert d["past_traj"].shape == (4, 11, 10, 2) and d["future_traj"].shape == (4, 11, 20, 2)


def test_header_is_plain_c(tmp_path):
    """include/sttode.h is a C ABI: it must compile as C99 (no C++ types, no default arguments), and a C translation unit that
    references every declared entry point must link against the library."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "sttode.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    src = tmp_path / "link_all.c"
    names = declared_symbols()
    src.write_text('#include "sttode.h"\n#include <stdio.h>\nint main(void) {\n  const void *p[] = {' + ", ".join(f"(const void *)&{n}" for n in names) +
                   '};\n  printf("%d %d\\n", (int)(sizeof p / sizeof p[0]), sttode_version());\n  return 0;\n}\n')
    exe = tmp_path / "link_all"
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir, "-lsttode_sm100",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == len(names) and int(out[1]) == 100


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` under a multi-rank launch: rank 0 alone runs and prints ONE JSON line carrying the contract's keys; the other
    ranks exit 0 without output (the driver launches the arm with torchrun for N > 1)."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env["RANK"] = env["LOCAL_RANK"] = "0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-scenes", "4"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "scenes/s" and d["n_gpus"] == 2 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and "workload" in d["config"]


def test_ctypes_signatures_match_header_prototypes():
    """Every prototype of include/sttode.h against its ctypes signature in sttode_b200/_lib.py: same number of parameters, and per position the same
    class (pointer / int / float / long / 64-bit unsigned = size_t and unsigned long long / unsigned / double), same for the return type.  ctypes
    would silently marshal a wrong signature; this is the check that the Python mirror calls the ABI the header declares."""
    import ctypes as C
    text = open(os.path.join(ROOT, "include", "sttode.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(sttode_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S)
    assert {n for _, n, _ in protos} == set(L.SIGNATURES)
    scalars = {"int": "int", "float": "float", "size_t": "u64", "unsigned long long": "u64", "long": "long", "double": "double", "unsigned": "uint"}

    def header_kind(decl, has_name):
        d = " ".join(decl.replace("const ", " ").split())
        if d in ("void", ""):
            return None
        if "*" in d:
            return "ptr"
        if has_name:
            d = d.rsplit(" ", 1)[0]
        return scalars[d]

    def ctypes_kind(t):
        if t is None:
            return None
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int: "int", C.c_float: "float", C.c_size_t: "u64", C.c_ulonglong: "u64", C.c_long: "long", C.c_double: "double", C.c_uint: "uint"}[t]

    for ret, name, params in protos:
        want = [k for k in (header_kind(p, True) for p in params.split(",")) if k is not None]
        res, args = L.SIGNATURES[name]
        assert [ctypes_kind(a) for a in args] == want, name
        assert ctypes_kind(res) == header_kind(ret, False), name


def test_ctypes_struct_fields_match_header_structs():
    """Field names, order and nesting of every struct in include/sttode.h against the ctypes mirror (sizes alone would not catch two swapped pointers)."""
    text = open(os.path.join(ROOT, "include", "sttode.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    mirror = {"sttode_mhgsa_weights": L.MhgsaWeights, "sttode_ode_weights": L.OdeWeights, "sttode_encoder_weights": L.EncoderWeights,
              "sttode_mlp3_weights": L.Mlp3Weights, "sttode_decompose_weights": L.DecomposeWeights, "sttode_model_weights": L.ModelWeights,
              "sttode_mhgsa_grads": L.MhgsaGrads, "sttode_ode_grads": L.OdeGrads, "sttode_mlp3_grads": L.Mlp3Grads, "sttode_decompose_grads": L.DecomposeGrads}
    structs = re.findall(r"typedef struct (sttode_[a-z0-9_]+) \{(.*?)\} \1;", text, flags=re.S)
    assert {n for n, _ in structs} == set(mirror)
    for name, body in structs:
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            first, *more = [d.strip() for d in decl.split(",")]
            toks = first.replace("*", " * ").split()
            base = " ".join(t for t in toks[:-1] if t not in ("const", "*"))
            for d in [toks[-1]] + [m.replace("*", "").strip() for m in more]:
                m = re.fullmatch(r"([a-z0-9_]+)(?:\[(\d+)\])?", d)
                fields.append((m.group(1), base if base.startswith("sttode_") else None, int(m.group(2)) if m.group(2) else None))
        got = mirror[name]._fields_
        assert [f[0] for f in got] == [f[0] for f in fields], name
        for (fname, nested, count), (_, ctype) in zip(fields, got):
            if nested:
                want = mirror[nested] * count if count else mirror[nested]
                assert ctype is want or (count and ctype._type_ is mirror[nested] and ctype._length_ == count), (name, fname)


def test_model_level_calls_have_no_cpu_fallback():
    """STTODENet on CPU tensors: inference, the batched forecast, the host-buffer evaluation and a train()-mode step all raise SttodeError (no device,
    no silent PyTorch path); argument errors of evaluate_host are reported before the library is touched."""
    from sttode_b200 import synth
    from sttode_b200.net import STTODENet
    m = STTODENet(make_args("eth"), torch.device("cpu")).eval()
    obs, gt = synth.eth_scene(1, 5)
    sp = torch.tensor([0, 5], dtype=torch.int32)
    m.set_data(None, obs, gt, torch.ones(5, 8), torch.ones(5, 12))
    for call in (lambda: m.inference(None), lambda: m.forecast_scenes(obs, sp), lambda: m.evaluate_host(obs, gt, sp), lambda: m.forward()):
        with pytest.raises(L.SttodeError):
            call()
    with pytest.raises(L.SttodeError, match=r"expected obs \[rows,2,8\]"):
        m.evaluate_host(obs[:, :, :5], gt, sp)
    n = STTODENet(make_args("nba", 10, 20), torch.device("cpu")).eval()
    d = synth.nba_batch(0, 3)
    with pytest.raises(L.SttodeError, match=r"expected obs \[B,N,10,2\]"):
        n.evaluate_host(d["past_traj"][:, :, :4], d["future_traj"])
    with pytest.raises(L.SttodeError):
        n.evaluate_host(d["past_traj"], d["future_traj"])
    n.set_data_nba(d)
    with pytest.raises(L.SttodeError):
        n.train().forward()


def test_product_never_touches_the_oracle_or_the_reference_tree():
    """The oracle is test infrastructure: no module of the package imports it (statically: AST of every file; dynamically: importing every module of the
    package leaves no `oracle*` entry in sys.modules), nothing under sttode_b200/ or include/ names /root/reference, and the C sources link nothing from oracle/."""
    import ast
    import glob
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "sttode_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            mods = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""] if isinstance(node, ast.ImportFrom) else []
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), path
    for path in glob.glob(os.path.join(pkg, "**", "*"), recursive=True) + glob.glob(os.path.join(ROOT, "include", "*")):
        if os.path.isfile(path) and not path.endswith((".so", ".pyc", ".srchash")):
            text = open(path, errors="replace").read()
            assert "/root/reference" not in text and "oracle/_ref" not in text, path
    code = ("import sys; sys.path.insert(0, %r); import sttode_b200, sttode_b200._lib, sttode_b200.net, sttode_b200.layers, sttode_b200.mhgsa, sttode_b200.pmath,"
            "sttode_b200.sampler, sttode_b200.parallel, sttode_b200.io, sttode_b200.synth, sttode_b200.compat, sttode_b200.build; sttode_b200._lib.load();"
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; print('BAD' if bad else 'CLEAN', bad)") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "CLEAN" in out.stdout, out.stdout + out.stderr
