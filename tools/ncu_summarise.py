"""Developer tool: `ncu -i rep --page raw --csv` -> the markdown table committed under profiles/ and traffic.json.
usage: ncu_summarise.py raw.csv out.md traffic.json "<source note>" """
import csv, json, sys
from collections import OrderedDict

def main(raw, out_md, out_json, note):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    def val(r, name, to=None):
        if name not in ix:
            name = [h for h in hdr if h.endswith("." + name)][0]
        v = float(r[ix[name]].replace(",", ""))
        u = units[ix[name]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
        return v * scale[u] if u in scale else v
    tensor = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    lines = [note, "", "| kernel | us | dram rd GB | dram wr GB | dram % | tensor pipe % | warps active % | regs | grid |", "|---|---|---|---|---|---|---|---|---|"]
    agg = OrderedDict()
    for r in data:
        name = r[ix["Kernel Name"]]
        us, rd, wr = val(r, "gpu__time_duration.sum"), val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        lines.append("| %s | %.3f | %.3f | %.3f | %.3f | %.3f | %.3f | %d | %d |" % (
            name[:36], us, rd, wr, val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), val(r, tensor),
            val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), val(r, "launch__registers_per_thread"), val(r, "launch__grid_size")))
        key = name.split("(")[0].replace("void ", "").split("<")[0].replace("tc::", "").replace("rt::", "")
        n, b = agg.get(key, (0, 0.0))
        agg[key] = (n + 1, b + (rd + wr) * 1e9)
    open(out_md, "w").write("\n".join(lines) + "\n")
    if out_json:
        old = json.load(open(out_json))
        old["source"] = note
        for k, (n, b) in agg.items():
            old["dram_bytes_per_launch"][k] = b / n
            old["launches_per_step"][k] = n
        json.dump(old, open(out_json, "w"), indent=1)
    print("\n".join(lines))

if __name__ == "__main__":
    main(*sys.argv[1:5])
