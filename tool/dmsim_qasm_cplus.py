#!/usr/bin/env python3
"""dmsim_qasm_cplus.py -- translates OpenQASM to a DM-Sim C++ driver (reference tool/dmsim_qasm_cplus.py).

Same command line (-i/--input, -o/--output, default dmsim_circuit.cpp) and the same statistics on stdout.  The emitted
program includes "dmsim_b200.hpp" instead of the reference's dmsim_cpu_omp.hpp and builds with
    g++ -O2 -std=c++17 -I include dmsim_circuit.cpp -L dm-sim_b200/lib -ldmsim_b200 -Wl,-rpath,dm-sim_b200/lib
--segment N splits prepare_circuit() into functions of N statements (circuit partitioner for very long circuits).
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(argv=None):
    parser = argparse.ArgumentParser(description="DM_Sim Assembler for OpenQASM-V2.0: translating OpenQASM to DM_sim "
                                                 "native simulation circuit code.")
    parser.add_argument("--input", "-i", required=True, help="input OpenQASM file, such as adder.qasm")
    parser.add_argument("--output", "-o", default="dmsim_circuit.cpp",
                        help="output DM_Sim circuit C++ file (default: dmsim_circuit.cpp)")
    parser.add_argument("--segment", type=int, default=4096, help="statements per prepare_circuit segment")
    args = parser.parse_args(argv)
    qasm = importlib.import_module("dm-sim_b200.qasm")
    with open(args.input) as f:
        text = f.read()
    cpp, stats = qasm.translate_cplus(text, segment=args.segment)
    with open(args.output, "w") as f:
        f.write(cpp)
    print("== DM-Sim: Translating " + args.input + " to " + args.output + " ==")
    print("Number of qubits: " + str(stats["n_qubits"]))
    print("Number of basic gates: " + str(stats["basic_gates"]))
    print("Number of cnot gates: " + str(stats["cnot_gates"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
