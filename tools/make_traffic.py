"""profiles/r2_traffic.json from ncu --set full reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of the
dominant kernels (bench.py copies the matching entry into roofline.traffic).
    python tools/make_traffic.py <step report (4096^2 Jacobi step)> <conv report 1024^2> <conv report 512^2> > profiles/r2_traffic.json
A conv launch is identified by its template (<3, 64, ...> = 3x3, Cout 64) and, among those, the longest one at the
report's resolution = the 128->64 full-resolution layer (Cin = 128 is the largest contraction with that template)."""
import csv, io, json, subprocess, sys


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = r[0], r[1], r[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def byt(d, key):
        v, u = float(d[ix[key]]), units[ix[key]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    out = []
    for d in data:
        out.append({"name": d[ix["Kernel Name"]], "us": float(d[ix["gpu__time_duration.sum"]]) *
                    {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1),
                    "bytes": byt(d, "dram__bytes_read.sum") + byt(d, "dram__bytes_write.sum")})
    return out


def main():
    step, conv1024, conv512 = sys.argv[1:4]
    t = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from round-2 ncu --set full captures "
                     "(profiles/r2_ncu_*_summary.txt); bench.py copies the matching entry into roofline.traffic"}
    jac = [r for r in rows_of(step) if "k_jacobi2d_blocked<" in r["name"] and
           any(tag in r["name"] for tag in ("<8, 0, 0,", "<12, 0, 0,", "(int)8, (bool)0, (bool)0", "(int)12, (bool)0, (bool)0"))]
    if jac:     # the middle launches of a solve: continued from p, no residual
        t["k_jacobi2d_blocked @4096x4096"] = int(sum(r["bytes"] for r in jac) / len(jac))
    # conv reports: exactly ONE MultiScaleNet forward (tools/prof_msnet.py, 16 tcgen05 launches in the order of
    # multi_scale_net.py:109-127: quarter 4, half 6, full 6); the full-resolution 64->128 / 128->64 layers are launches 12 / 13
    for rep, res in ((conv1024, 1024), (conv512, 512)):
        if rep == "-":
            continue
        c = [r for r in rows_of(rep) if "k_conv_tc" in r["name"]]
        if len(c) >= 16:
            c = c[-16:]
            t[f"k_conv_tc 64->128 k3 @{res}x{res}"] = int(c[12]["bytes"])
            t[f"k_conv_tc 128->64 k3 @{res}x{res}"] = int(c[13]["bytes"])
    print(json.dumps(t, indent=2))


if __name__ == "__main__":
    main()
