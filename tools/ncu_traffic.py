"""profiles/r02_ncu_traffic.json out of a full ncu capture of the substep kernel (tools/gpu_final.sh): DRAM bytes per
launch and a few figures next to them, stamped with the hash of the library the capture ran (bench.py quotes
roofline.traffic only when that hash is the one of the library it runs).
usage: python tools/ncu_traffic.py gpurun_out/r02_full_k_substep_resident.ncu-rep gpurun_out/r02_bench_under_ncu_full.log"""
import csv, importlib, io, json, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rep, log = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))


def num(key, scale=None):
    v = float(d[key])
    unit = u[key]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3}.get(unit, 1)
    return v * mult


line = [l for l in open(log) if l.startswith("{")][-1]
bench = json.loads(line)
out = {
    "workload": bench["config"]["workload"],
    "kernel": d["Kernel Name"],
    "lib_sha16": importlib.import_module("soft-body-simulator_b200.build").source_id(),  # sources unchanged since the capture
    "dram_bytes_read": int(num("dram__bytes_read.sum")),
    "dram_bytes_write": int(num("dram__bytes_write.sum")),
    "gpu_time_us_under_ncu": num("gpu__time_duration.sum"),
    "registers_per_thread": int(float(d["launch__registers_per_thread"])),
    "warp_instructions": int(float(d["smsp__inst_executed.sum"])),
    "shared_bank_conflicts": int(float(d["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"])),
    "warps_active_pct_of_peak": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
    "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
    "how": "ncu --set full --cache-control none --clock-control none -k regex:k_substep_resident -s 25 -c 1 python bench.py "
           "--steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-sub (tools/gpu_final.sh); one launch = one substep",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
# the raw page of that one launch, and the stall reasons of the source page summed over the kernel
open(os.path.join(ROOT, "profiles", "r02_final_ncu_full_k_substep_resident_raw.csv"), "w").write(raw)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cats = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {c: 0.0 for c in cats}
for r in rows[2:]:
    for c in cats:
        try:
            tot[c] += float(r[ix[c]])
        except (ValueError, IndexError):
            pass
sel = tot.get("stall_selected", 1.0) or 1.0
with open(os.path.join(ROOT, "profiles", "r02_final_ncu_stalls_k_substep_resident.txt"), "w") as f:
    f.write("warp stall reasons of %s (one launch = one substep of config 3), ncu --set full, samples summed over the SASS of the "
            "kernel, relative to 'selected' (= issued)\n" % d["Kernel Name"][:60])
    for c, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write("%-28s %8.0f samples  %6.3f\n" % (c[6:], v, v / sel))

