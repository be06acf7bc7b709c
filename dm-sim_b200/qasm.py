"""OpenQASM-2 front-end of the engine: a Python-3 rewrite of the reference's translator tool/dmsim_qasm.py.

Two consumers of one parser:

* ``translate(text)``  -> the DM-Sim Python script the reference tool writes (same header, same
  ``sim.append(sim.<OP>(params, qubits))`` lines, user gates as ``def`` functions, same footer) plus the statistics
  it prints (qubits / basic gates / cnot gates, from the same GATE_TABLE / CX_TABLE, reference :31-87).
* ``load(text)``       -> ``(n_qubits, gates)`` with gates as ``(OP, [qubits], theta, phi, lam)`` tuples placed in the
  Gate fields exactly as the reference's static factories do (src/dmsim_nvgpu_omp.cuh:580-767), ready for
  ``Simulation`` / ``pack_gates`` -- no script round trip needed.

Semantics kept from the reference (tool/dmsim_qasm.py:233-330): qregs are flattened into one global qubit numbering
in declaration order; ``measure / barrier / OPENQASM / include / creg / if / reset`` are dropped; a gate applied to
whole registers is broadcast; parameter expressions are evaluated with ``pi``.  Beyond the reference: parameters of
user-defined gates may be used inside their bodies, and several statements may share a line.
"""
from __future__ import annotations

import ast
import math
import re

# number of "basic gates" / CNOTs each op counts for in the printed statistics (reference :31-87, kept verbatim,
# including its under-counts for cu3 / c3x / c3sqrtx / c4x)
STANDARD_GATE_TABLE = {"u3": 1, "u2": 1, "u1": 1, "cx": 1, "id": 1, "x": 1, "y": 1, "z": 1, "h": 1, "s": 1, "sdg": 1,
                       "t": 1, "tdg": 1, "rx": 1, "ry": 1, "rz": 1, "c1": 1, "c2": 1}
COMPOSITION_GATE_TABLE = {"cz": 3, "cy": 3, "swap": 3, "ch": 11, "ccx": 15, "cswap": 17, "crx": 5, "cry": 4, "crz": 4,
                          "cu1": 5, "cu3": 5, "rxx": 7, "rzz": 3, "rccx": 9, "rc3x": 18, "c3x": 27, "c3sqrtx": 27,
                          "c4x": 87}
GATE_TABLE = {**STANDARD_GATE_TABLE, **COMPOSITION_GATE_TABLE}
CX_TABLE = {"u3": 0, "u2": 0, "u1": 0, "cx": 1, "id": 0, "x": 0, "y": 0, "z": 0, "h": 0, "s": 0, "sdg": 0, "t": 0,
            "tdg": 0, "rx": 0, "ry": 0, "rz": 0, "c1": 0, "c2": 1, "cz": 1, "cy": 1, "swap": 3, "ch": 2, "ccx": 6,
            "cswap": 8, "crx": 2, "cry": 2, "crz": 2, "cu1": 2, "cu3": 2, "rxx": 2, "rzz": 2, "rccx": 3, "rc3x": 6,
            "c3x": 6, "c3sqrtx": 6, "c4x": 18}
# ops of enum OP that qelib1.inc does not define but the engine accepts as lower-case names too
EXTRA_OPS = {"ryy": 7, "r": 1, "srn": 1, "w": 1}
OTHER_KEYS = ("measure", "barrier", "OPENQASM", "include", "creg", "if", "reset", "opaque")

# (number of parameters, number of qubits, where the parameters land in Gate(theta, phi, lambda))
_SIG = {
    "u3": (3, 1, ("theta", "phi", "lam")), "u2": (2, 1, ("phi", "lam")), "u1": (1, 1, ("lam",)),
    "cx": (0, 2, ()), "id": (0, 1, ()), "x": (0, 1, ()), "y": (0, 1, ()), "z": (0, 1, ()), "h": (0, 1, ()),
    "s": (0, 1, ()), "sdg": (0, 1, ()), "t": (0, 1, ()), "tdg": (0, 1, ()),
    "rx": (1, 1, ("theta",)), "ry": (1, 1, ("theta",)), "rz": (1, 1, ("phi",)),
    "cz": (0, 2, ()), "cy": (0, 2, ()), "swap": (0, 2, ()), "ch": (0, 2, ()), "ccx": (0, 3, ()), "cswap": (0, 3, ()),
    "crx": (1, 2, ("lam",)), "cry": (1, 2, ("lam",)), "crz": (1, 2, ("lam",)), "cu1": (1, 2, ("lam",)),
    "cu3": (3, 2, ("theta", "phi", "lam")), "rxx": (1, 2, ("theta",)), "rzz": (1, 2, ("theta",)),
    "rccx": (0, 3, ()), "rc3x": (0, 4, ()), "c3x": (0, 4, ()), "c3sqrtx": (0, 4, ()), "c4x": (0, 5, ()),
    "ryy": (1, 2, ("theta",)), "r": (1, 1, ("theta",)), "srn": (0, 1, ()), "w": (0, 1, ()),
}


class QasmError(ValueError):
    pass


_SAFE = {"pi": math.pi, "sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "ln": math.log,
         "sqrt": math.sqrt, "asin": math.asin, "acos": math.acos, "atan": math.atan}


_BINOPS = {ast.Add: lambda a, b: a + b, ast.Sub: lambda a, b: a - b, ast.Mult: lambda a, b: a * b,
           ast.Div: lambda a, b: a / b, ast.Pow: lambda a, b: a ** b}


def _eval_node(node, names):
    """Arithmetic over numbers, the names in `names` and calls of the functions in it -- nothing else (the text comes from
    an untrusted .qasm file: no attribute access, subscripts, comprehensions, lambdas ...)."""
    if isinstance(node, ast.Expression):
        return _eval_node(node.body, names)
    if isinstance(node, ast.Constant) and type(node.value) in (int, float):
        return node.value
    if isinstance(node, ast.Name) and node.id in names and not callable(names[node.id]):
        return names[node.id]
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.UAdd, ast.USub)):
        v = _eval_node(node.operand, names)
        return -v if isinstance(node.op, ast.USub) else +v
    if isinstance(node, ast.BinOp) and type(node.op) in _BINOPS:
        return _BINOPS[type(node.op)](_eval_node(node.left, names), _eval_node(node.right, names))
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and callable(names.get(node.func.id)) and not node.keywords:
        return names[node.func.id](*[_eval_node(a, names) for a in node.args])
    raise ValueError(f"unsupported syntax: {ast.dump(node)[:60]}")


def _eval_raw(expr: str, env=None):
    """the Python value of a parameter expression (int stays int: the reference prints str(eval(expr)))"""
    try:
        v = _eval_node(ast.parse(expr.strip(), mode="eval"), {**_SAFE, **(env or {})})
        float(v)
        return v
    except Exception as e:  # noqa: BLE001
        raise QasmError(f"cannot evaluate parameter expression {expr!r}: {e}") from None


def _eval(expr: str, env=None):
    return float(_eval_raw(expr, env))


def _split_args(s: str):
    """split on commas that are not inside parentheses / brackets"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


_HEAD = re.compile(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*", re.S)


def _split_stmt(s: str):
    """(name, parameter text or None, rest) of `name(params) rest`; the parameter list is cut at the MATCHING parenthesis,
    so that nested ones (`u1(-(pi/4)) q[0]`) parse -- the reference's non-greedy regex stops at the first `)`."""
    m = _HEAD.match(s)
    if not m:
        return None
    name, i = m.group(1), m.end()
    par = None
    if i < len(s) and s[i] == "(":
        depth = 0
        for j in range(i, len(s)):
            depth += s[j] == "("
            depth -= s[j] == ")"
            if depth == 0:
                par, i = s[i + 1:j], j + 1
                break
        else:
            return None
    return name, par, s[i:].strip()


class Program:
    def __init__(self):
        self.qregs = {}        # name -> (start, length)
        self.n_qubits = 0
        self.user_gates = {}   # name -> (params, qargs, [(op, param_exprs, qarg_names)])
        self.items = []        # ("comment", text) | ("gatedef", name) | ("call", op, [param exprs], [qubit args])
        self.gate_num = 0
        self.cx_num = 0
        self.user_counts = {}  # name -> (basic gates, cx)


def parse(text: str) -> Program:
    prog = Program()
    # comments: whole-line comments are kept (the reference copies them into the script), trailing ones dropped
    lines = text.splitlines()
    stream = []
    for ln in lines:
        st = ln.strip()
        if st.startswith("//"):
            stream.append(("comment", st[2:]))
            continue
        if "//" in st:
            st = st[:st.index("//")]
        if st:
            stream.append(("code", st))
    # glue code lines and split into statements on ';' while keeping gate bodies {...} together
    i = 0
    buf = ""

    def flush_statement(s):
        s = s.strip()
        if not s:
            return
        m = _split_stmt(s)
        if not m:
            raise QasmError(f"cannot parse statement {s!r}")
        op, par, rest = m
        if op in OTHER_KEYS:
            return
        if op == "qreg":
            mm = re.match(r"([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*(\d+)\s*\]", rest)
            if not mm:
                raise QasmError(f"bad qreg declaration {s!r}")
            prog.qregs[mm.group(1)] = (prog.n_qubits, int(mm.group(2)))
            prog.n_qubits += int(mm.group(2))
            return
        params = _split_args(par) if par else []
        qargs = _split_args(rest)
        low = op.lower()
        if op in prog.user_gates:
            n, cx = prog.user_counts[op]
            reps = _broadcast(prog, qargs)
            prog.gate_num += n * len(reps)
            prog.cx_num += cx * len(reps)
            prog.items.append(("call", op, params, qargs))
        elif low in GATE_TABLE or low in EXTRA_OPS:
            if low in ("c1", "c2"):
                raise QasmError("c1 / c2 take a matrix and cannot be written in OpenQASM; use Simulation.C1 / C2")
            reps = _broadcast(prog, qargs)
            prog.gate_num += GATE_TABLE.get(low, EXTRA_OPS.get(low, 1)) * len(reps)
            prog.cx_num += CX_TABLE.get(low, 0) * len(reps)
            prog.items.append(("call", low, params, qargs))
        else:
            raise QasmError(f"Unknown symbol: {op}")

    while i < len(stream):
        kind, st = stream[i]
        i += 1
        if kind == "comment":
            if not buf.strip():
                prog.items.append(("comment", st))
            continue
        buf += " " + st
        while True:
            b = buf.lstrip()
            if b.startswith("gate ") or b.startswith("gate\t"):
                if "}" not in b:
                    break  # need more lines
                head, body = b[:b.index("{")], b[b.index("{") + 1:b.index("}")]
                buf = b[b.index("}") + 1:]
                _define_gate(prog, head, body)
                continue
            if ";" not in b:
                break
            stmt, buf = b[:b.index(";")], b[b.index(";") + 1:]
            flush_statement(stmt)
    if buf.strip():
        flush_statement(buf)
    return prog


def _define_gate(prog: Program, head: str, body: str):
    m = re.match(r"gate\s+([A-Za-z_][A-Za-z0-9_]*)\s*(?:\((.*?)\))?\s*(.*)$", head.strip(), re.S)
    if not m:
        raise QasmError(f"bad gate definition {head!r}")
    name = m.group(1)
    params = [p for p in _split_args(m.group(2) or "") if p]
    qargs = [q for q in _split_args(m.group(3)) if q]
    ops, n, cx = [], 0, 0
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        mm = _split_stmt(stmt)
        if not mm:
            raise QasmError(f"cannot parse statement {stmt!r} in a gate body")
        op, par, rest = mm
        if op == "barrier":
            continue
        low = op.lower()
        if op in prog.user_gates:
            n += prog.user_counts[op][0]
            cx += prog.user_counts[op][1]
            ops.append((op, _split_args(par) if par else [], _split_args(rest)))
        elif low in GATE_TABLE or low in EXTRA_OPS:
            n += GATE_TABLE.get(low, EXTRA_OPS.get(low, 1))
            cx += CX_TABLE.get(low, 0)
            ops.append((low, _split_args(par) if par else [], _split_args(rest)))
        else:
            raise QasmError(f'"{stmt}" is not a gate in Function!')
    prog.user_gates[name] = (params, qargs, ops)
    prog.user_counts[name] = (n, cx)
    prog.items.append(("gatedef", name))


def _qubit(prog: Program, arg: str):
    """'q[3]' -> global index; 'q' -> list of global indices (whole register)"""
    m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\s*(\d+)\s*\])?$", arg.strip())
    if not m or m.group(1) not in prog.qregs:
        raise QasmError(f"unknown quantum register in {arg!r}")
    start, length = prog.qregs[m.group(1)]
    if m.group(2) is None:
        return [start + i for i in range(length)]
    idx = int(m.group(2))
    if idx >= length:
        raise QasmError(f"qubit index out of range in {arg!r}")
    return start + idx


def _broadcast(prog: Program, qargs):
    """list of qubit tuples, one per broadcast repetition (reference paramlist_to_ga :138-174)"""
    res = [_qubit(prog, a) for a in qargs]
    widths = [len(r) for r in res if isinstance(r, list) and len(r) > 1]
    n = widths[0] if widths else 1
    out = []
    for b in range(n):
        tup = []
        for r in res:
            if isinstance(r, list):
                if len(r) == 1:
                    tup.append(r[0])
                elif b < len(r):
                    tup.append(r[b])
                else:
                    raise QasmError("Error in Syntax!")
            else:
                tup.append(r)
        out.append(tuple(tup))
    return out


# ------------------------------------------------------------------------------------------------------------------
def _emit(gates, prog, op, pvals, qubits):
    """one built-in op with numeric parameters on concrete qubits -> gate tuple in factory field order"""
    npar, nq, fields = _SIG[op]
    if len(pvals) != npar or len(qubits) != nq:
        raise QasmError(f"{op}: expected {npar} parameter(s) and {nq} qubit(s)")
    f = {"theta": 0.0, "phi": 0.0, "lam": 0.0}
    for name, v in zip(fields, pvals):
        f[name] = v
    gates.append((op.upper(), [int(q) for q in qubits], f["theta"], f["phi"], f["lam"]))


def _expand_user(gates, prog, name, pvals, qubits, depth=0):
    if depth > 64:
        raise QasmError("gate definitions nest too deeply")
    params, qargs, ops = prog.user_gates[name]
    if len(pvals) != len(params) or len(qubits) != len(qargs):
        raise QasmError(f"{name}: wrong number of arguments")
    env = dict(zip(params, pvals))
    qmap = dict(zip(qargs, qubits))
    for op, pex, qa in ops:
        vals = [_eval(e, env) for e in pex]
        qs = []
        for a in qa:
            if a not in qmap:
                raise QasmError(f"{name}: unknown qubit argument {a!r}")
            qs.append(qmap[a])
        if op in prog.user_gates:
            _expand_user(gates, prog, op, vals, qs, depth + 1)
        else:
            _emit(gates, prog, op, vals, qs)


def load(text: str):
    """(n_qubits, gates): the circuit as gate tuples (OP, [qubits], theta, phi, lam)."""
    prog = parse(text)
    gates = []
    for it in prog.items:
        if it[0] != "call":
            continue
        _, op, pex, qargs = it
        vals = [_eval(e) for e in pex]
        for qubits in _broadcast(prog, qargs):
            if op in prog.user_gates:
                _expand_user(gates, prog, op, vals, list(qubits))
            else:
                _emit(gates, prog, op, vals, list(qubits))
    return prog.n_qubits, gates


def load_file(path: str):
    with open(path) as f:
        return load(f.read())


def _fmt_param(expr: str, symbols=()):
    """numeric text of a parameter as the reference prints it (str(float)); symbolic inside gate definitions"""
    try:
        return str(_eval_raw(expr))
    except QasmError:
        if any(re.search(r"\b%s\b" % re.escape(s), expr) for s in symbols):
            return expr
        raise


def translate(text: str, module: str = "dmsim_py_omp_wrapper"):
    """Returns (script_text, stats) where stats = dict(n_qubits, basic_gates, cnot_gates)."""
    prog = parse(text)
    out = ["import sys\n", f"import {module} as dmsim\n\n", "if (len(sys.argv) != 3):\n",
           "\tprint('$python circuit.py n_qubits n_gpus')\n", "\texit()\n\n",
           "sim = dmsim.Simulation(int(sys.argv[1]), int(sys.argv[2]))\n\n"]
    uses_pi = False
    for it in prog.items:
        if it[0] == "comment":
            out.append("#" + it[1] + "\n")
        elif it[0] == "gatedef":
            name = it[1]
            params, qargs, ops = prog.user_gates[name]
            s = "def " + name + "(sim" + "".join(", " + p for p in params) + "".join(", " + q for q in qargs) + "):\n"
            for op, pex, qa in ops:
                ptxt = [_fmt_param(e, params) for e in pex]
                if any("pi" in p for p in ptxt):
                    uses_pi = True
                args = ", ".join(ptxt + list(qa))
                if op in prog.user_gates:
                    s += "\t" + op + "(sim, " + args + ")\n"
                else:
                    s += "\tsim.append(sim." + op.upper() + "(" + args + "))\n"
            if not ops:
                s += "\tpass\n"
            out.append(s + "\n")
        else:
            _, op, pex, qargs = it
            ptxt = [str(_eval_raw(e)) for e in pex]
            for qubits in _broadcast(prog, qargs):
                args = ", ".join(ptxt + [str(q) for q in qubits])
                if op in prog.user_gates:
                    out.append(op + "(sim, " + args + ")\n")
                else:
                    out.append("sim.append(sim." + op.upper() + "(" + args + "))\n")
    if uses_pi:
        out.insert(2, "from math import pi\n")
    out.append("\nsim.upload()\nsim.run()\nsim.measure(10)\n")
    stats = {"n_qubits": prog.n_qubits, "basic_gates": prog.gate_num, "cnot_gates": prog.cx_num}
    return "".join(out), stats


# ------------------------------------------------------------------------------------------------------------------
# C++ emitter: the reference's tool/dmsim_qasm_cplus.py (OpenQASM -> a C++ driver on the Simulation API, :308-382) for
# include/dmsim_b200.hpp.  Same structure (user gates become functions, the main circuit becomes prepare_circuit(),
# main() = upload / sim / measure(5) / print_measurement), with three repairs: gate functions get typed parameters (the
# reference emits untyped ones, :286-291), every factory object is deleted after append() copied it (:331-344), and
# long circuits are PARTITIONED into segments of `segment` statements (the paper's "Partitioner": one huge function
# takes the C++ compiler minutes).
def translate_cplus(text: str, header: str = "dmsim_b200.hpp", segment: int = 4096):
    """Returns (cpp_text, stats) where stats = dict(n_qubits, basic_gates, cnot_gates, segments)."""
    prog = parse(text)
    out = ["#include <stdio.h>\n", f'#include "{header}"\n', "//Use the DMSim namespace to enable C++/CUDA APIs\n",
           "using namespace DMSim;\n\n",
           "// append() deep-copies the gate: the factory object is released right away\n",
           "#define DMSIM_APPEND(G) do { Gate* g_ = Simulation::G; sim.append(g_); delete g_; } while (0)\n\n"]
    body = []  # statements of the main circuit
    for it in prog.items:
        if it[0] == "comment":
            (body if body else out).append("//" + it[1] + "\n")
        elif it[0] == "gatedef":
            name = it[1]
            params, qargs, ops = prog.user_gates[name]
            s = "void " + name + "(Simulation &sim" + "".join(", const ValType " + p for p in params) + \
                "".join(", const IdxType " + q for q in qargs) + ")\n{\n"
            for op, pex, qa in ops:
                ptxt = [_fmt_param(e, params) for e in pex]
                ptxt = [re.sub(r"\bpi\b", "PI", p) for p in ptxt]
                args = ", ".join(ptxt + list(qa))
                if op in prog.user_gates:
                    s += "\t" + op + "(sim, " + args + ");\n"
                else:
                    s += "\tDMSIM_APPEND(" + op.upper() + "(" + args + "));\n"
            out.append(s + "}\n\n")
        else:
            _, op, pex, qargs = it
            ptxt = [repr(float(_eval_raw(e))) for e in pex]
            for qubits in _broadcast(prog, qargs):
                args = ", ".join(ptxt + [str(q) for q in qubits])
                if op in prog.user_gates:
                    body.append("\t" + op + "(sim, " + args + ");\n")
                else:
                    body.append("\tDMSIM_APPEND(" + op.upper() + "(" + args + "));\n")
    segment = max(1, int(segment))
    n_seg = max(1, (len(body) + segment - 1) // segment)
    if n_seg == 1:
        out.append("void prepare_circuit(Simulation &sim)\n{\n" + "".join(body) + "}\n\n")
    else:
        for k in range(n_seg):
            out.append(f"static void prepare_circuit_{k}(Simulation &sim)\n{{\n" + "".join(body[k * segment:(k + 1) * segment]) + "}\n\n")
        out.append("void prepare_circuit(Simulation &sim)\n{\n" + "".join(f"\tprepare_circuit_{k}(sim);\n" for k in range(n_seg)) + "}\n\n")
    out.append("int main()\n{\n\tsrand(RAND_SEED);\n\tint n_qubits=" + str(prog.n_qubits) + ";\n\tint n_gpus=1;\n"
               "\tSimulation sim(n_qubits, n_gpus);\n\tprepare_circuit(sim);\n\tsim.upload();\n\tsim.sim();\n"
               "\tauto* res = sim.measure(5);\n\tprint_measurement(res, n_qubits, 5);\n\tdelete[] res;\n\treturn 0;\n}\n")
    stats = {"n_qubits": prog.n_qubits, "basic_gates": prog.gate_num, "cnot_gates": prog.cx_num, "segments": n_seg}
    return "".join(out), stats
