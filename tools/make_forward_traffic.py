"""Turns an ncu capture of K1 at 4K (gpurun_out/k1_full.ncu-rep, `ncu --set full` of tools/perf_forward.py) into
profiles/forward_traffic.json: DRAM bytes and warp-instructions per launch, stamped with the digest of the kernel source the
library was built from. bench.py quotes the file only while the digest matches the tree (a stale capture is refused).
usage: python tools/make_forward_traffic.py gpurun_out/k1_full.ncu-rep "<one-line description of the capture>" """
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
rows = [r for r in rows[2:] if "forward_kernel" in r[hdr.index("Kernel Name")]]
col = lambda name: [float(r[hdr.index(name)]) for r in rows]
unit = lambda name: list(csv.reader(out.splitlines()))[1][hdr.index(name)]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = sum(col("dram__bytes_read.sum")) / len(rows) * scale[unit("dram__bytes_read.sum")]
wr = sum(col("dram__bytes_write.sum")) / len(rows) * scale[unit("dram__bytes_write.sum")]
inst = sum(col("smsp__inst_executed.sum")) / len(rows)
dur = sum(col("gpu__time_duration.sum")) / len(rows)
j = {"kernel": rows[0][hdr.index("Kernel Name")], "launches_averaged": len(rows), "frame": "3840x2160",
     "dram_bytes_per_launch": round(rd + wr), "dram_bytes_read": round(rd), "dram_bytes_written": round(wr),
     "warp_instructions_per_launch": round(inst), "duration_under_ncu": f"{dur:.1f} {unit('gpu__time_duration.sum')}",
     "source_digest": bench._forward_source_digest(), "capture": f"{os.path.basename(rep)} (ncu --set full, tools/perf_forward.py); {note}"}
json.dump(j, open(os.path.join(ROOT, "profiles", "forward_traffic.json"), "w"), indent=1)
print(json.dumps(j, indent=1))
