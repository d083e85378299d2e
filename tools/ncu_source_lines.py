"""Dynamic instruction counts per CUDA source line for one kernel of an ncu report: joins the per-SASS
`Instructions Executed` column of `ncu --page source` with nvdisasm's line info of the matching cubin.
    python tools/ncu_source_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled-name substring> <cells>"""
import collections, csv, io, re, subprocess, sys
rep, kre, cubin, key, cells = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
dis = subprocess.run(["nvdisasm", "-c", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
st = [i for i, l in enumerate(dis) if l.startswith(".text.") and key in l][0]
en = next((i for i in range(st + 1, len(dis)) if dis[i].startswith(".text.")), len(dis))
cur, lm = None, {}
for l in dis[st:en]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    a = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if a:
        lm[int(a.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
ix = {n: i for i, n in enumerate(rows[1])}
end = next((i for i in range(2, len(rows)) if rows[i][:2] == ["Address", "Source"]), len(rows))
sec = rows[2:end]
base = int(sec[0][0], 16)
cnt, tot = collections.Counter(), 0.0
for r in sec:
    try:
        off, ex = int(r[0], 16) - base, float(r[ix["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    cnt[lm.get(off)] += ex
    tot += ex
print(f"{rows[0][1][:90]}\nwarp instructions: {tot:.0f}  = {tot * 32 / cells:.0f} thread-instructions per cell")
byf = collections.Counter()
for k, v in cnt.items():
    byf[k[0] if k else None] += v
print("by file:", {a: f"{100 * b / tot:.1f}%" for a, b in byf.most_common()})
for k, v in cnt.most_common(25):
    print(f"  {100 * v / tot:5.1f}%  {k}")
