"""CPU: the OpenQASM front-end (dm-sim_b200/qasm.py, tool/dmsim_qasm.py) -- against the reference's own translator
(executed under a Python-3 shim when its tree is present), the golden gate list, and the oracle."""
import glob
import importlib
import io
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
qasm = importlib.import_module("dm-sim_b200.qasm")

SMALL = """OPENQASM 2.0;
include "qelib1.inc";
// a comment line
qreg a[2];
qreg b[3];
creg c[5];
gate foo(theta, phi) x, y { u3(theta, phi, pi/2) x; cx x, y; rz(-theta/2) y; }
h a;            // broadcast over the register
cx a, b[1];
foo(pi/3, 0.25) a[1], b[2];
cu1(pi/4) b[0], b[2]; barrier a; measure a[0] -> c[0];
ccx a[0], a[1], b[0];
"""


def test_load_semantics():
    n, g = qasm.load(SMALL)
    assert n == 5
    names = [x[0] for x in g]
    assert names == ["H", "H", "CX", "CX", "U3", "CX", "RZ", "CU1", "CCX"]
    assert g[2][1] == [0, 3] and g[3][1] == [1, 3]          # b[1] is global qubit 3
    assert g[4][1] == [1] and abs(g[4][2] - np.pi / 3) < 1e-15 and g[4][3] == 0.25 and abs(g[4][4] - np.pi / 2) < 1e-15
    assert g[6][0] == "RZ" and abs(g[6][3] + np.pi / 6) < 1e-15 and g[6][2] == 0.0   # RZ stores its angle in phi
    assert g[7][0] == "CU1" and abs(g[7][4] - np.pi / 4) < 1e-15                     # CU1 stores it in lambda
    with pytest.raises(qasm.QasmError):
        qasm.load("qreg q[2]; frobnicate q[0];")
    with pytest.raises(qasm.QasmError):
        qasm.load("qreg q[2]; h q[5];")


def test_translated_script_shape_and_stats():
    script, stats = qasm.translate(SMALL)
    assert script.startswith("import sys\nimport dmsim_py_omp_wrapper as dmsim\n")
    assert "sim = dmsim.Simulation(int(sys.argv[1]), int(sys.argv[2]))" in script
    assert "sim.append(sim.H(0))\nsim.append(sim.H(1))\n" in script
    assert "def foo(sim, theta, phi, x, y):" in script
    assert script.endswith("\nsim.upload()\nsim.run()\nsim.measure(10)\n")
    assert stats == {"n_qubits": 5, "basic_gates": 2 + 2 + 3 + 5 + 15, "cnot_gates": 2 + 1 + 2 + 6}
    compile(script, "circuit.py", "exec")  # valid Python 3


def test_golden_vqe_gate_list_matches_parser():
    if not os.path.exists(REF):
        pytest.skip("reference tree not present")
    z = np.load(os.path.join(ROOT, "tests", "golden", "vqe_uccsd_n8.npz"))
    dm = importlib.import_module("dm-sim_b200")
    n, g = qasm.load_file(os.path.join(REF, "benchmark", "vqe_uccsd_n8.qasm"))
    rec, _ = dm.pack_gates(g)
    assert n == 8 and len(g) == 10808 and rec.tobytes() == z["gates"].tobytes()


def _run_reference_translator(path, out):
    """Executes the reference's Python-2 tool under Python 3 with its two dict-concatenation lines patched in memory."""
    with open(os.path.join(REF, "tool", "dmsim_qasm.py")) as f:
        src = f.read()
    src = src.replace("dict(STANDARD_GATE_TABLE.items() + COMPOSITION_GATE_TABLE.items())",
                      "dict(list(STANDARD_GATE_TABLE.items()) + list(COMPOSITION_GATE_TABLE.items()))")
    src = src.replace("dict(STANDARD_CX_TABLE.items() + COMPOSITION_CX_TABLE.items())",
                      "dict(list(STANDARD_CX_TABLE.items()) + list(COMPOSITION_CX_TABLE.items()))")
    argv, stdout = sys.argv, sys.stdout
    sys.argv, sys.stdout = ["dmsim_qasm.py", "-i", path, "-o", out], io.StringIO()
    try:
        exec(compile(src, "ref_dmsim_qasm.py", "exec"), {"__name__": "__main__"})
        printed = sys.stdout.getvalue()
    finally:
        sys.argv, sys.stdout = argv, stdout
    with open(out) as f:
        return f.read(), printed


@pytest.mark.parametrize("name", ["bv_n15", "qft_n15", "adder_n9", "cc_n15", "vqe_uccsd_n8", "qec_n5", "w_state_n3",
                                  "grover_n3", "sat_n10", "deutsch_n5"])
def test_same_output_as_the_reference_translator(tmp_path, name):
    if not os.path.exists(REF):
        pytest.skip("reference tree not present")
    path = os.path.join(REF, "benchmark", name + ".qasm")
    ref_script, ref_printed = _run_reference_translator(path, str(tmp_path / "ref.py"))
    with open(path) as f:
        mine, stats = qasm.translate(f.read())

    def body(s):  # the gate-appending lines, whitespace-normalised
        return [ln.strip() for ln in s.splitlines() if "sim.append" in ln or ln.strip().endswith(")") and "(sim" in ln]
    assert body(mine) == body(ref_script)
    assert f"Number of qubits: {stats['n_qubits']}" in ref_printed
    assert f"Number of basic gates: {stats['basic_gates']}" in ref_printed
    assert f"Number of cnot gates: {stats['cnot_gates']}" in ref_printed


def test_cli(tmp_path):
    src = tmp_path / "c.qasm"
    src.write_text(SMALL)
    out = tmp_path / "c.py"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tool", "dmsim_qasm.py"), "-i", str(src), "-o", str(out)],
                       capture_output=True, text=True, check=True)
    assert "Number of qubits: 5" in r.stdout and "Number of basic gates: 27" in r.stdout
    assert "sim.append(sim.CCX(0, 1, 2))" in out.read_text()


def test_generators_match_the_benchmark_files():
    if not os.path.exists(REF):
        pytest.skip("reference tree not present")
    C = importlib.import_module("dm-sim_b200.circuits")
    assert qasm.load_file(os.path.join(REF, "benchmark", "qft_n15.qasm")) == (15, C.qft(15))
    assert qasm.load_file(os.path.join(REF, "benchmark", "bv_n15.qasm")) == (15, C.bv(15))
    n, g = qasm.load_file(os.path.join(REF, "benchmark", "adder_n9.qasm"))
    assert n == 10 and len(g) == 30 and sorted(x[0] for x in g) == sorted(x[0] for x in C.adder_n10())


def test_qasm_circuit_against_oracle(dm, oracle_mod):
    import kernel_emulator as ke
    n, g = qasm.load(SMALL)
    re, im = oracle_mod.Oracle(n).sim(g).dm()
    v0 = np.zeros(4 ** n, dtype=np.complex128); v0[0] = 1
    out = ke.run_plan_dev(dm.plan_json(n, 1, g), v0)
    assert np.abs(out - (re + 1j * im).reshape(-1)).max() < 1e-12


def test_cplus_translator_emits_a_buildable_driver(tmp_path):
    """tool/dmsim_qasm_cplus.py (reference tool/dmsim_qasm_cplus.py:308-382): user gates become typed functions, the
    main circuit prepare_circuit() -- split into segments when long --, same statistics; the program compiles and links
    against the drop-in header."""
    cpp, stats = qasm.translate_cplus(SMALL, segment=3)
    assert stats["n_qubits"] == 5 and stats["segments"] == 3
    _, py_stats = qasm.translate(SMALL)
    assert (stats["basic_gates"], stats["cnot_gates"]) == (py_stats["basic_gates"], py_stats["cnot_gates"])
    assert "void foo(Simulation &sim, const ValType theta, const ValType phi, const IdxType x, const IdxType y)" in cpp
    assert "DMSIM_APPEND(U3(theta, phi, 1.5707963267948966, x));" in cpp and "DMSIM_APPEND(RZ(-theta/2, y));" in cpp
    assert "\tfoo(sim, 1.0471975511965976, 0.25, 1, 4);" in cpp
    assert cpp.count("DMSIM_APPEND(H(") == 2 and "DMSIM_APPEND(CX(1, 3));" in cpp       # broadcast over the register
    assert "prepare_circuit_2(sim);" in cpp and "int n_qubits=5;" in cpp
    src = tmp_path / "c.cpp"
    src.write_text(cpp)
    lib = os.path.join(ROOT, "dm-sim_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o",
                    str(tmp_path / "c"), "-L", lib, "-ldmsim_b200", "-Wl,-rpath," + lib], check=True)
    out = tmp_path / "cli.cpp"
    q = tmp_path / "in.qasm"
    q.write_text(SMALL)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tool", "dmsim_qasm_cplus.py"), "-i", str(q), "-o", str(out)],
                       capture_output=True, text=True, check=True)
    assert "Number of qubits: 5" in r.stdout and "Number of basic gates: " + str(stats["basic_gates"]) in r.stdout
    assert "void prepare_circuit(Simulation &sim)" in out.read_text()


def _run_reference_cplus_translator(path, out):
    """The reference's tool/dmsim_qasm_cplus.py executed under Python 3 (same in-memory patches as above)."""
    with open(os.path.join(REF, "tool", "dmsim_qasm_cplus.py")) as f:
        src = f.read()
    src = src.replace("dict(STANDARD_GATE_TABLE.items() + COMPOSITION_GATE_TABLE.items())",
                      "dict(list(STANDARD_GATE_TABLE.items()) + list(COMPOSITION_GATE_TABLE.items()))")
    src = src.replace("dict(STANDARD_CX_TABLE.items() + COMPOSITION_CX_TABLE.items())",
                      "dict(list(STANDARD_CX_TABLE.items()) + list(COMPOSITION_CX_TABLE.items()))")
    argv, stdout = sys.argv, sys.stdout
    sys.argv, sys.stdout = ["dmsim_qasm_cplus.py", "-i", path, "-o", out], io.StringIO()
    try:
        exec(compile(src, "ref_dmsim_qasm_cplus.py", "exec"), {"__name__": "__main__"})
        printed = sys.stdout.getvalue()
    finally:
        sys.argv, sys.stdout = argv, stdout
    with open(out) as f:
        return f.read(), printed


@pytest.mark.parametrize("name", ["bv_n15", "qft_n15", "adder_n9", "cc_n15", "qec_n5", "w_state_n3", "grover_n3", "sat_n10"])
def test_cplus_translator_against_the_reference(tmp_path, name):
    """Same gate statements (factory name + arguments, in order) and the same statistics as the reference's
    tool/dmsim_qasm_cplus.py for the main circuit of the benchmark files."""
    import re
    if not os.path.exists(REF):
        pytest.skip("reference tree not present")
    path = os.path.join(REF, "benchmark", name + ".qasm")
    ref_cpp, ref_printed = _run_reference_cplus_translator(path, str(tmp_path / "ref.cpp"))
    with open(path) as f:
        mine, stats = qasm.translate_cplus(f.read())

    def calls(s, pat):  # (NAME, [numeric args]) of every appended gate
        out = []
        for m in re.finditer(pat, s):
            try:
                out.append((m.group(1), [float(a) for a in m.group(2).split(",")]))
            except ValueError:     # symbolic arguments inside a user-defined gate: compare as text
                out.append((m.group(1), [a.strip() for a in m.group(2).split(",")]))
        return out
    ref_calls = calls(ref_cpp, r"sim\.append\(Simulation::([A-Z0-9]+)\(([^)]*)\)\)")
    my_calls = calls(mine, r"DMSIM_APPEND\(([A-Z0-9]+)\(([^)]*)\)\)")
    assert len(my_calls) == len(ref_calls) and len(my_calls) > 0
    for (a, pa), (b, pb) in zip(my_calls, ref_calls):
        assert a == b
        if all(isinstance(x, float) for x in pa + pb):
            assert np.allclose(pa, pb, rtol=0, atol=1e-12)
    assert f"Number of qubits: {stats['n_qubits']}" in ref_printed
    assert f"Number of basic gates: {stats['basic_gates']}" in ref_printed
    assert f"Number of cnot gates: {stats['cnot_gates']}" in ref_printed


def test_parameter_expressions_are_not_python_eval():
    """Parameter text comes from an untrusted file: arithmetic, pi and the qelib functions only; nested parentheses in a
    parameter list parse (the reference's non-greedy regex stops at the first ')')."""
    qasm = importlib.import_module("dm-sim_b200.qasm")
    src = ('OPENQASM 2.0;\ninclude "qelib1.inc";\ngate foo(a,b) x { u1((a-b)/2) x; }\nqreg q[2];\n'
           'u1(-(pi/4)) q[0];\nfoo(pi, sin(pi/2)) q[1];\n')
    n, gates = qasm.load(src)
    assert n == 2 and [g[0] for g in gates] == ["U1", "U1"]
    lam = [g[4] if g[4] else g[2] for g in gates]
    assert abs(abs(lam[0]) - np.pi / 4) < 1e-15 and abs(abs(lam[1]) - (np.pi - 1) / 2) < 1e-15
    for bad in ("().__class__", "__import__('os')", "[1][0]", "pi.real", "(lambda: 1)()"):
        with pytest.raises(qasm.QasmError):
            qasm._eval_raw(bad)
