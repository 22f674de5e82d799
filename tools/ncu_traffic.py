"""Sum the DRAM bytes of every k_extend launch of an ncu --csv metrics log and divide by the rays the program reports:
    python tools/ncu_traffic.py <ncu csv log> <program stdout with the stats json> -> one json object for profiles/extend_traffic.json"""
import csv, json, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
i_name, i_metric, i_unit, i_val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
n = 0
for r in rows[1:]:
    if "k_extend" not in r[i_name] or r[i_metric] not in tot:
        continue
    tot[r[i_metric]] += float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
    n += r[i_metric] == "gpu__time_duration.sum"
m = re.search(r'\{"ray_segments".*\}', open(sys.argv[2], errors="replace").read())
st = json.loads(m.group(0))
rays = st["ray_segments"]
print(json.dumps({"kernel": "k_extend<3,false,0>", "launches": n, "rays": rays, "dram_read_bytes": tot["dram__bytes_read.sum"], "dram_write_bytes": tot["dram__bytes_write.sum"],
                  "dram_bytes_per_ray": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / rays, "extend_ms_under_ncu": tot["gpu__time_duration.sum"]}))
