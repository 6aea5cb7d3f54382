"""profiles/traffic.json from ncu launch lists (CSV written by
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X <cmd>).

   python scripts/ncu_traffic.py <train_launches.csv> [<decode_n1.csv> <decode_n2.csv> <n2 - n1>]

(the two decode lists are prof_decode.py runs with n1 and n2 new tokens: their difference isolates n2 - n1 decode steps
from the shared prefill)

Per kernel family: launches, total / average duration, DRAM bytes read + written per launch.  bench.py reads the
gemm_bf16_kernel entry for roofline.traffic and the decode_step entry for roofline_decode.traffic."""
import csv, json, os, re, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    i0 = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[i0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    ui = hdr.index("Metric Unit")
    per = defaultdict(dict)
    names = {}
    for r in rows[i0 + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = r[ui]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(unit, 1.0)
        per[r[ii]][r[mi]] = v * scale
        names[r[ii]] = r[ki]
    return per, names


def family(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"^ergm::", "", name)
    return re.split(r"[<(]", name)[0]


def summarise(per, names):
    fam = defaultdict(lambda: dict(launches=0, us=0.0, dram=0.0))
    for k, m in per.items():
        f = fam[family(names[k])]
        f["launches"] += 1
        f["us"] += m.get("gpu__time_duration.sum", 0.0)
        f["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    return fam


def main():
    train = sys.argv[1]
    per, names = parse(train)
    fam = summarise(per, names)
    total_us = sum(f["us"] for f in fam.values())
    out = {}
    print("%-34s %8s %10s %8s %14s" % ("kernel family", "launches", "total us", "share", "DRAM B/launch"))
    for n, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        print("%-34s %8d %10.1f %7.1f%% %14.0f" % (n, f["launches"], f["us"], 100 * f["us"] / total_us, f["dram"] / f["launches"]))
        out[n] = {"launches": f["launches"], "total_us": round(f["us"], 1), "share": round(f["us"] / total_us, 4),
                  "dram_bytes_per_launch": round(f["dram"] / f["launches"]), "source": os.path.basename(train)}
    g = [f for n, f in fam.items() if n.startswith("gemm")]
    if g:
        out["gemm_bf16_kernel"] = {"launches": sum(f["launches"] for f in g),
                                   "dram_bytes_per_launch": round(sum(f["dram"] for f in g) / sum(f["launches"] for f in g)),
                                   "share": round(sum(f["us"] for f in g) / total_us, 4),
                                   "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the GEMM launches of "
                                             + os.path.basename(train)}
    if len(sys.argv) > 4:
        steps = int(sys.argv[4])
        tot = []
        for path in sys.argv[2:4]:
            dper, dnames = parse(path)
            tot.append((sum(m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0) for m in dper.values()),
                        len(dper)))
        dram, nk = tot[1][0] - tot[0][0], tot[1][1] - tot[0][1]
        out["decode_step"] = {"dram_bytes_per_step": round(dram / steps), "kernels_per_step": nk / steps,
                              "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, difference of %s and %s over %d "
                                        "decode steps" % (os.path.basename(sys.argv[3]), os.path.basename(sys.argv[2]), steps)}
        print("decode: %.1f MB of DRAM traffic per step, %.0f kernels per step" % (dram / steps / 1e6, nk / steps))
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
