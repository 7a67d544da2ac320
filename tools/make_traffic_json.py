"""profiles/onet_decode_traffic.json (read by bench.py for roofline.traffic) from the committed ncu raw CSV of the decoder
capture of the same round:  python tools/make_traffic_json.py profiles/<tag>_onet_decode_raw.csv <objects>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, objects = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rows = list(csv.reader(open(path)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def get(name):
    v, u = float(vals[col[name]].replace(",", "")), units[col[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0}.get(u, 1.0)
    return v * scale


rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
out = {"kernel": "onet_decode_kernel", "objects": objects, "points_per_object": 32768,
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
       "algorithmic_bytes_per_launch": objects * 32768 * 4 + 32768 * 12 + objects * 23552 + 1310720,
       "kernel_ms_under_ncu": get("gpu__time_duration.sum"),
       "tensor_pipe_active_pct_of_nominal": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
       if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
       "note": "logits are only partly written back to DRAM inside the kernel window: the 126 MB L2 still holds the tail",
       "source": f"ncu --set full --clock-control none -k regex:onet_decode -s 1 -c 1 python tools/prof_decoder.py {objects} 1 "
                 f"-> {os.path.relpath(path, ROOT)} (tools/run_round_profiles.sh)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "onet_decode_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
