"""Join an ncu SASS-level source page (csv) with nvdisasm line info to get executed
warp-instructions per CUDA source line.  Usage:
  sass_by_line.py <ncu_source.csv> <nvdisasm_with_lineinfo.dis> <mangled kernel substr> <frames>"""
import csv, re, sys
src_csv, dis, kname, frames = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
# 1. instruction -> (file, line) from nvdisasm
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l)
cur = ('?', 0); insts = []
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//-----'):
        if insts: break
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = m.group(3)
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        insts.append((int(m.group(1), 16), m.group(2), cur))
# 2. executed counts from ncu (same order)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; body = rows[2:]
ia, isrc, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed')
ist = hdr.index('Warp Stall Sampling (All Samples)')
iw = hdr.index('L1 Wavefronts Shared')
assert len(body) == len(insts), (len(body), len(insts))
agg = {}
tot = 0; totw = 0
for r, (addr, txt, loc) in zip(body, insts):
    ex = int(r[iex]); st = int(r[ist] or 0); w = int(r[iw] or 0)
    a = agg.setdefault(loc, [0, 0, 0]); a[0] += ex; a[1] += st; a[2] += w
    tot += ex; totw += w
print(f"total warp-inst/frame {tot/frames:.1f}  smem wavefronts/frame {totw/frames:.1f}")
src = {}
for loc in agg:
    f = loc[0]
    if f not in src:
        import glob
        c = glob.glob(f'/root/repo/**/{f}', recursive=True)
        src[f] = open(c[0]).read().split('\n') if c else []
allst = sum(a[1] for a in agg.values())
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 60]:
    s = src[loc[0]][loc[1] - 1].strip()[:80] if src[loc[0]] and loc[1] - 1 < len(src[loc[0]]) else ''
    print(f"{a[0]/frames:8.1f} inst  {100*a[1]/allst:5.1f}% stall  {a[2]/frames:7.1f} wf  {loc[0]}:{loc[1]}: {s}")
