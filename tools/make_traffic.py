"""profiles/traffic.json from one `ncu --set full` capture of the scoring kernel (raw CSV page):
dram__bytes_read.sum + dram__bytes_write.sum per launch -- bench.py copies it into roofline.traffic."""
import csv, json, os, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(h)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
r = next(r for r in rows[2:] if "deeplab_score_vec4_kernel" in r[ix["Kernel Name"]])
tot = sum(float(r[ix[k]].replace(",", "")) * scale[units[ix[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
out = {"deeplab_score_vec4_kernel": tot, "grid": r[ix["Grid Size"]], "source": f"ncu --set full --clock-control none, {os.path.basename(sys.argv[1])}, 16x19x1024x2048 launch (tools/prof_run.py score)",
       "tag": sys.argv[2] if len(sys.argv) > 2 else ""}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "gpurun_out", "traffic.json"), "w"), indent=1)
print(out)
