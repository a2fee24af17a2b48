#!/usr/bin/env python
"""Turn one `ncu --set full` capture (.ncu-rep) into the records kept under profiles/: the details page as text, the raw page
as csv, and an entry of profiles/pgs_traffic.json (measured DRAM bytes of the solver kernel, keyed on workload, batch and the
hash of the solver sources so that bench.py only uses it while it describes the kernels it runs).

    python tools/ncu_summary.py gpurun_out/r2c/k_pgs_pipe.ncu-rep --name r02_k_pgs_pipe_mb1k --workload marble_box_1k --scenes 2664 --kernel k_pgs_pipe
"""
import argparse, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--name", required=True); ap.add_argument("--workload", required=True)
ap.add_argument("--scenes", type=int, required=True); ap.add_argument("--kernel", required=True)
ap.add_argument("--command", default="SLRBS_GRAPH=0 bench.py --steps 2 --warmup 3")
a = ap.parse_args()

det = subprocess.run(["ncu", "-i", a.rep, "--page", "details"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", a.name + "_ncu_details.txt"), "w").write(det)
open(os.path.join(ROOT, "profiles", a.name + "_ncu_raw.csv"), "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(name, scale_units=True):
    u, v = col[name]
    x = float(v.replace(",", ""))
    if scale_units:
        x *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9}.get(u, 1.0)
    return x


import bench
rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
rec = {"workload": a.workload, "scenes": a.scenes, "kernel": a.kernel, "source_hash": bench.solver_source_hash(),
       "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
       "ncu_duration_ms": num("gpu__time_duration.sum") * 1e3,
       "shared_wavefronts_per_launch": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False),
       "shared_bank_conflict_wavefronts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", False),
       "registers": num("launch__registers_per_thread", False),
       "source": f"profiles/{a.name}_ncu_raw.csv (ncu --set full, one launch, {a.command})"}
p = os.path.join(ROOT, "profiles", "pgs_traffic.json")
allrec = json.load(open(p)) if os.path.exists(p) else []
allrec = [r for r in allrec if not (r.get("workload") == a.workload and r.get("scenes") == a.scenes)] + [rec]
json.dump(allrec, open(p, "w"), indent=1)
print(json.dumps(rec, indent=1))
