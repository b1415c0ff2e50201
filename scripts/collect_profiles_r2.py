"""Round 2 (profiles/r02_*).  Turns the files scripts/gpu_r2_profile.sh left in gpurun_out/ into the committed evidence under profiles/:
    python scripts/collect_profiles_r2.py a
per-launch time + DRAM table, roofline_traffic.json (read by bench.py), ncu --set full summaries, bench lines, logs."""
import csv, json, os, shutil, subprocess, sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

lines = [l for l in open(os.path.join(G, "launches_%s.csv" % tag)) if l.startswith('"')]
by, order = {}, []
for r in csv.DictReader(lines):
    i = r["ID"]
    if i not in by:
        by[i] = {"name": r["Kernel Name"], "grid": r["Grid Size"]}
        order.append(i)
    by[i][r["Metric Name"]] = float(r["Metric Value"])
    by[i][r["Metric Name"] + "_u"] = r["Metric Unit"]
ids = [i for i in order if "mel_to_act" in by[i]["name"]]
fw = order[order.index(ids[1]):order.index(ids[2])]
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot_t = tot_b = tc_b = tc_t = tp_w = tp_all = 0.0
TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
have_tp = all(TP in by[i] for i in fw)
out = []
for k, i in enumerate(fw):
    d = by[i]
    t = d["gpu__time_duration.sum"] / 1e3
    rb = d["dram__bytes_read.sum"] * mult[d["dram__bytes_read.sum_u"]]
    wb = d["dram__bytes_write.sum"] * mult[d["dram__bytes_write.sum_u"]]
    n = d["name"].split("(")[0].split("::")[-1][:30]
    line = "%3d %-30s grid=%-14s %8.1f us  dram rd %7.1f MB  wr %7.1f MB" % (k, n, d["grid"], t, rb / 1e6, wb / 1e6)
    if have_tp:
        ghz = d.get("sm__cycles_elapsed.avg.per_second", 0.0)
        line += "  tensor pipe %5.1f %% @ %.2f GHz" % (d[TP], ghz * 1e-9 if ghz > 1e6 else ghz)
        tp_all += d[TP] * t
    out.append(line)
    tot_t += t
    tot_b += rb + wb
    if "tc_kernel" in n or "tz_kernel" in n:   # conv_tc / pair_tc / rb_tc / pair_tz
        tc_b += rb + wb
        tc_t += t
        if have_tp:
            tp_w += d[TP] * t
hdr = ("forward #1 of `python bench.py --steps 1 --warmup 3 --passes 1 --quick --no-side` under ncu (serialised, cold-cache, no "
       "programmatic overlap: compare shares): %d launches, %.1f us, DRAM %.1f MB; tcgen05 launches: %.1f us, DRAM %.1f MB"
       % (len(fw), tot_t, tot_b / 1e6, tc_t, tc_b / 1e6))
if have_tp:
    hdr += ("; tensor pipe active (sm__pipe_tensor_cycles_active, time-weighted): %.1f %% over the tcgen05 launches, %.1f %% "
            "over the whole forward" % (tp_w / tc_t, tp_all / tot_t))
name = "r02_launches_%s_time_dram.txt" % tag
open(os.path.join(P, name), "w").write(hdr + "\n" + "\n".join(out) + "\n")
json.dump({"dram_bytes_per_step_tcgen05": tc_b, "dram_bytes_per_step_all": tot_b,
           "tensor_pipe_active_pct_tcgen05": (tp_w / tc_t) if have_tp else None,
           "tensor_pipe_active_pct_forward": (tp_all / tot_t) if have_tp else None,
           "source": "profiles/%s (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, one forward)" % name},
          open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
print(hdr)
for src, dst in (("pair_s1k11", "pair_s1k11"), ("rb_s2k3", "rb_s2k3"), ("tz_s3k11", "tz_s3k11"), ("mel", "mel")):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (src, tag))
    if os.path.exists(rep):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, "14"],
                             capture_output=True, text=True).stdout
        open(os.path.join(P, "r02_ncu_full_%s_%s.txt" % (dst, tag)), "w").write(txt)
for src, dst in (("bench_%s.json", "r02_bench_%s.json"), ("bench_ref_%s.json", "r02_bench_reference_%s.json"),
                 ("bench_fp16_%s.json", "r02_bench_fp16_operands_%s.json"), ("pytest_gpu_%s.log", "r02_pytest_gpu_%s.log"),
                 ("parity_margins_%s.jsonl", "r02_parity_margins_%s.jsonl")):
    if os.path.exists(os.path.join(G, src % tag)):
        shutil.copy(os.path.join(G, src % tag), os.path.join(P, dst % tag))
