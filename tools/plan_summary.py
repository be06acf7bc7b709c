"""Print the schedule the planner/encoder produce for a workload: per sweep the tile bits, groups, rounds and the
device op codes of each round (host only, no GPU).  Usage: python tools/plan_summary.py qft_n15 [world]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dm = importlib.import_module("dm-sim_b200")
NAMES = {0: "D1", 2: "MONO1", 3: "SRN", 4: "D2", 6: "PERM2", 7: "DIAGR", 8: "RR", 9: "RI", 10: "STAR", 11: "HAD", 12: "DIAGP", 13: "CP2", 14: "QFT2"}
n, gates = bench.workload(sys.argv[1])
p = dm.plan_json(n, int(sys.argv[2]) if len(sys.argv) > 2 else 1, gates)
print({k: v for k, v in p.items() if not isinstance(v, (list, dict))})
for s in p["steps"]:
    if s["kind"] != "sweep":
        print(s["kind"]); continue
    d = s["dev"]
    print(f"sweep k={s['k']} in_pos={s['in_pos']} oop={s['out_of_place']} groups={len(d['groups'])} rounds={len(d['rounds'])} ops={len(d['ops'])} stars={len(d['stars'])}")
    offs = {o["off"]: i for i, o in enumerate(d["ops"])}
    for r in d["rounds"]:
        i0 = offs.get(r["first"], 0)
        ops = d["ops"][i0:i0 + r["count"]]
        print("   round n_iter", r["n_iter"], " ".join(f"{NAMES.get(o['code'], o['code'])}{o['pos'] if o['code'] not in (7, 10) else ''}" + (f"[{bin(o["aux"] & 15).count("1")}]" if o["code"] in (10, 11) else (f"[{8 - bin(o['aux'] & 255).count('1')}]" if o["code"] == 12 else (f"[{16 - bin(o['aux'] & 0xffff).count('1')}]" if o["code"] == 7 else ""))) for o in ops))
