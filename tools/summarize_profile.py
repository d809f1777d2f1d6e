"""Turn gpurun_out/*.ncu-rep + launch list into the tracked summaries under profiles/.

  python tools/summarize_profile.py <tag> <full.ncu-rep> <launches.csv> <clips in the captured run>
"""
import csv, gzip, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, launches, clips = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
summary = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    k = {"kernel": d.get("Kernel Name"), "clips_in_capture": clips}
    for key in KEYS:
        if key in d:
            try:
                k[key] = float(d[key])
            except ValueError:
                k[key] = d[key]
            k[key + "__unit"] = units[hdr.index(key)]
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                v = float(d[h])
            except ValueError:
                continue
            if v >= 0.05:
                k.setdefault("stall_per_issue", {})[h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = round(v, 3)
    summary.append(k)
with open(os.path.join(out_dir, f"{tag}_ncu_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)

sys.path.insert(0, ROOT)
import importlib.util
_spec = importlib.util.spec_from_file_location("bench_for_hash", os.path.join(ROOT, "bench.py"))
_bench = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_bench)
SOURCE_HASH = _bench.kernel_source_hash()
# the shipped default-plan kernel: frames_fast_2048<16, false, true, /*TM*/true, kMelUnrDefault>
KSUB = "frames_fast_2048ILi16ELb0ELb1ELb1ELi235340547E"      # bench.py quotes a capture only when this matches the built sources

def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

for k in summary:
    if k["kernel"] and "frames_fast" in k["kernel"]:
        rd = k["dram__bytes_read.sum"] * unit_scale(k["dram__bytes_read.sum__unit"])
        wr = k["dram__bytes_write.sum"] * unit_scale(k["dram__bytes_write.sum__unit"])
        with open(os.path.join(out_dir, "frames_fast_traffic.json"), "w") as f:
            json.dump({"source": f"profiles/{tag}_ncu_summary.json (ncu --set full, one launch)",
                       "source_hash": SOURCE_HASH, "clips_in_capture": clips, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "dram_bytes_per_launch": rd + wr, "dram_bytes_per_clip": (rd + wr) / clips}, f, indent=1)
        with open(os.path.join(out_dir, "frames_fast_pipes.json"), "w") as f:
            json.dump({"source": f"profiles/{tag}_ncu_summary.json (ncu --set full, one launch, {clips} clips)",
                       "source_hash": SOURCE_HASH,
                       "fp32_pipe_cycles_active_pct": k.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                       "fma_pipe_pct": k.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                       "alu_pipe_pct": k.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                       "lsu_pipe_pct": k.get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                       "issue_slots_busy_pct": k.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "shared_mem_pipe_pct": k.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                       "dram_pct": k.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                       "registers_per_thread": k.get("launch__registers_per_thread"),
                       "warp_instructions_per_frame": k.get("smsp__inst_executed.sum", 0) / (clips * 130.0)}, f, indent=1)
        break

# launch list: keep our kernels' rows
keep = []
with open(launches) as f:
    for line in f:
        if line.startswith('"ID"') or "hlmc::" in line:
            keep.append(line)
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.writelines(keep)

# per-source-line instruction / stall / shared-wavefront profile of the frames kernel
src_csv = f"/tmp/{tag}_src.csv"
with open(src_csv, "w") as f:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the report may hold several kernels: keep the first frames_fast section
    parts = txt.split('"Kernel Name",')
    sec = next(p for p in parts if p.startswith('"void hlmc::frames_fast'))
    f.write('"Kernel Name",' + sec)
lib = os.path.join(ROOT, "hybrid_language_music_clustering_vae_b200", "libhlmc_b200.so")
tmp = f"/tmp/{tag}_cub"
os.makedirs(tmp, exist_ok=True)
subprocess.run(f"cd {tmp} && rm -f *.cubin && cuobjdump -xelf all {lib} > /dev/null && "
               f"nvdisasm --print-line-info -c hlmc_kernels.sm_100a.cubin > k.dis 2>/dev/null", shell=True, check=True)
frames = clips * 130
res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_by_line.py"), src_csv, f"{tmp}/k.dis",
                      KSUB, str(frames), "80"], capture_output=True, text=True)
with open(os.path.join(out_dir, f"{tag}_frames_fast_by_line.txt"), "w") as f:
    f.write("# executed warp-instructions, share of stall samples and shared-memory wavefronts per frame,\n"
            "# by CUDA source line (ncu source page joined with nvdisasm line info)\n" + res.stdout + res.stderr[-2000:])
# SASS listing of the frames kernel
sass = subprocess.run(f"cuobjdump -sass {lib}", shell=True, capture_output=True, text=True).stdout
chunks = sass.split("\tFunction : ")
pick = [c for c in chunks if c.startswith("_ZN4hlmc16" + KSUB)] + [c for c in chunks if c.startswith("_ZN4hlmc6db_dctILi40")]
with gzip.open(os.path.join(out_dir, f"{tag}_sass_frames_fast_db_dct.txt.gz"), "wt") as f:
    f.write("\n\tFunction : ".join(pick))
mix = {}
for c in pick[:1]:
    for line in c.splitlines():
        p = line.split("*/")
        if len(p) >= 2 and p[0].strip().startswith("/*"):
            op = p[1].strip().split()[0] if p[1].strip() else ""
            if op.startswith("@"):
                op = p[1].strip().split()[1]
            op = op.split(".")[0].rstrip(";")
            if op:
                mix[op] = mix.get(op, 0) + 1
with open(os.path.join(out_dir, f"{tag}_sass_mix_frames_fast.json"), "w") as f:
    json.dump(dict(sorted(mix.items(), key=lambda kv: -kv[1])), f, indent=1)
# the Blackwell / TMA evidence lines of the frames kernel
ev = []
for line in pick[0].splitlines() if pick else []:
    if any(op in line for op in ("UBLKCP", "SYNCS", "MUFU.SQRT", "FENCE.VIEW.ASYNC", "UTMA", "LDTM", "STTM", "UTCATOMSWS")):
        ev.append(line.rstrip())
packed = {op: [l.rstrip() for l in (pick[0].splitlines() if pick else []) if f" {op} " in l] for op in ("FFMA2", "FADD2", "FMUL2")}
with open(os.path.join(out_dir, f"{tag}_sass_evidence.txt"), "w") as f:
    f.write("# frames_fast_2048<16,0,1,TM,default steps>: TMA bulk copy (UBLKCP), mbarrier (SYNCS), async-proxy fence, MUFU lines,\n"
            "# Tensor Memory: allocation (UTCATOMSWS), table fill (STTM), table reads (LDTM = tcgen05.ld)\n")
    f.write("\n".join(ev) + "\n")
    f.write("# Blackwell packed FP32 (sm_100 only): " + ", ".join(f"{op} x{len(v)}" for op, v in packed.items()) +
            " static instructions; first lines of each (operand swizzles .LO_HI / sign patterns .NP fold the +-i rotations):\n")
    for op, v in packed.items():
        f.write("\n".join(v[:4]) + "\n")
print("wrote profiles/", tag)
