"""profiles/ncu_summary.json from an ncu capture of the traversal kernels — the per-ray figures bench.py quotes next to its
live roofline (lanes per instruction, warp instructions per ray, DRAM bytes per ray).

usage: python tools/ncu_to_json.py <capture.ncu-rep> <workload, e.g. C2> <counts.txt from BN_DEBUG_COUNTS=1> <first k_traverse launch captured (ncu -s)> "<source note>"

The capture lists k_traverse launches in stream order: extend b0, shadow b0, extend b1, shadow b1, ... of the first wave, so
launch k of the capture is (extend if (s+k) even else shadow) of bounce (s+k)//2; the rays of each launch come from the
bn_counts line of wave 0.  Every captured launch is kept; `extend` / `shadow` at the top are the bounce-1 launches (the
figures round 1 quoted), `by_bounce` holds the rest (the incoherent deep bounces)."""
import csv, json, os, re, subprocess, sys

rep, workload, counts_path, first, source = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
col = {name: i for i, name in enumerate(hdr)}
line = next(l for l in open(counts_path) if "wave 0" in l)
ext = [int(x) for x in re.findall(r"extend=(\d+)", line)]
sh = [int(x) for x in re.findall(r"shadow=(\d+)", line)]
def num(r, name):
    v = r[col[name]].replace(",", "") if name in col else ""
    return float(v) if v not in ("", "n/a") else None
units = rows[1]
def gb(r, name):   # ncu prints bytes in a unit of its choosing
    v, u = num(r, name), units[col[name]].lower()
    return None if v is None else v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
out = {"source": source, "by_bounce": {}}
for k, r in enumerate(rows[2:]):
    idx = first + k
    kind, bounce = ("extend" if idx % 2 == 0 else "shadow"), idx // 2
    rays = (ext if kind == "extend" else sh)[bounce]
    inst = num(r, "smsp__inst_executed.sum")
    dram = (gb(r, "dram__bytes_read.sum") or 0) + (gb(r, "dram__bytes_write.sum") or 0)
    e = {"kernel": r[col["Kernel Name"]][:48], "bounce": bounce, "rays": rays, "duration_ms_under_ncu": num(r, "gpu__time_duration.sum"),
         "lanes_per_inst": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), "warp_inst_per_ray": inst / rays if inst else None,
         "dram_bytes_per_ray": dram / rays, "dram_bytes_per_launch": dram,
         "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "l1tex_throughput_pct": num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
         "lts_throughput_pct": num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
         "l1_hit_pct": num(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
         "registers": num(r, "launch__registers_per_thread"), "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")}
    out["by_bounce"][f"{kind}_b{bounce}"] = e
    if bounce == 1 or kind not in out:
        out[kind] = e
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json")
allj = json.load(open(path)) if os.path.exists(path) else {}
allj[workload] = out
json.dump(allj, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
