"""profiles/ncu_traffic_<tag>.json from an `ncu --set full` capture of tools/profile_kernels.py: DRAM bytes
(read + write) and duration per hot kernel, grouped by the kernel families bench.py reports."""
import csv
import json
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rep = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/prof_r1d_kernels.ncu-rep"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    return float(r[idx[k]].replace(",", ""))


fam = {}
pending_attn_gemms = 0   # the two batched GEMMs xf_attn_bwd issues right after its key-stationary pass belong to that call
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("xf::", "").strip()
    unit_r, unit_w = rows[1][idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_write.sum"]]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    b = num(r, "dram__bytes_read.sum") * scale.get(unit_r, 1.0) + num(r, "dram__bytes_write.sum") * scale.get(unit_w, 1.0)
    us = num(r, "gpu__time_duration.sum") * {"usecond": 1.0, "us": 1.0, "nsecond": 1e-3, "ns": 1e-3, "msecond": 1e3}.get(
        rows[1][idx["gpu__time_duration.sum"]], 1.0)
    f = "gemm" if name.startswith("gemm") else "attn_fwd" if name.startswith("attn_fwd") else "attn_bwd" if name.startswith("attn_bwd") \
        else "layernorm_fwd" if name.startswith("layernorm_fwd") else "layernorm_bwd" if name.startswith("layernorm_bwd") else name
    if f == "attn_bwd" and "<3," in r[idx["Kernel Name"]]:
        pending_attn_gemms = 2
    elif f == "gemm" and pending_attn_gemms > 0:
        f, pending_attn_gemms = "attn_bwd", pending_attn_gemms - 1
        name += " (batched dQ / dK GEMM of xf_attn_bwd)"
    fam.setdefault(f, []).append({"kernel": name, "dram_bytes": b, "duration_us": us})
res = {"source": f"ncu --set full, tools/profile_kernels.py (Ego4Dv2 level-0 shape, B=13), profiles/ncu_full_{tag}.md", "families": {}}
for f, ks in fam.items():
    if f == "attn_bwd":
        res["families"][f] = {"traffic_bytes_per_call": sum(k["dram_bytes"] for k in ks), "kernels": ks,
                              "note": "one xf_attn_bwd call = key-stationary pass (dV + E to the scratch) + the batched dQ and dK GEMMs"}
    else:
        res["families"][f] = {"traffic_bytes_per_launch": ks[0]["dram_bytes"], "kernels": ks}
with open(f"profiles/ncu_traffic_{tag}.json", "w") as fo:
    json.dump(res, fo, indent=1)
print("wrote", f"profiles/ncu_traffic_{tag}.json", {f: round(v.get("traffic_bytes_per_call", v.get("traffic_bytes_per_launch")) / 1e6, 1) for f, v in res["families"].items()})
