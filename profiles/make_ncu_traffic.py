"""profiles/ncu_traffic.json from an extract_ncu.py digest: DRAM bytes per launch of the roofline kernels, stamped with the
hash of caliscope_b200/csrc/ the capture was taken from (bench.py only reports `traffic` when the hash still matches).
usage: python profiles/make_ncu_traffic.py profiles/r02/ncu_summary.jsonl cfg4"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import kernel_source_hash  # noqa: E402

digest, workload = sys.argv[1], sys.argv[2]
out_path = ROOT / "profiles" / "ncu_traffic.json"
data = json.loads(out_path.read_text()) if out_path.exists() else {}
per = {}
for line in open(digest):
    r = json.loads(line)
    name = r["kernel"].split("<")[0].split("::")[-1].replace("void ", "").strip()
    if "dram_read_MB" not in r:
        continue
    # keep the largest launch of each kernel (the predicated-off launches move nothing)
    tot = (r["dram_read_MB"] + r.get("dram_write_MB", 0.0)) * 1e6
    if tot > per.get(name, {}).get("dram_bytes_per_launch", -1):
        per[name] = {"dram_bytes_per_launch": tot, "time_us": r.get("time_us"), "csrc_sha16": kernel_source_hash(),
                     "dram_read_MB": r["dram_read_MB"], "dram_write_MB": r.get("dram_write_MB", 0.0)}
data[workload] = per
out_path.write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")
print(json.dumps(per, indent=1))
