"""profiles/traffic.json from an `ncu --set full` report: DRAM bytes per launch of a kernel + the hash of the source that was profiled.
usage: python devtools/update_traffic.py <report.ncu-rep> <kernel key> <chains> <source file> <sha256 of that source when profiled> <capture name>"""
import csv, io, json, os, subprocess, sys
rep, key, chains, src, sha, name = sys.argv[1:7]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(metric):
    i = hdr.index(metric)
    v = float(vals[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
tj = json.load(open(path)) if os.path.exists(path) else {}
tj[key] = dict(chains=int(chains), dram_bytes_per_launch=rd + wr, dram_bytes_read=rd, dram_bytes_write=wr, source=src, source_sha256=sha, capture=name)
json.dump(tj, open(path, "w"), indent=1)
print(tj[key])
