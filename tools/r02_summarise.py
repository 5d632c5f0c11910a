"""Turn the round-2 capture (tools/gpu_round2_capture.sh, files gpurun_out/<tag>_*) into the tracked summaries under profiles/.
Runs without a GPU (ncu -i reads the reports).    python tools/r02_summarise.py r2z"""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = sys.argv[1] if len(sys.argv) > 1 else "r2z"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def write(name, text):
    with open(os.path.join(P, name), "w") as f:
        f.write(text)
    print("wrote profiles/" + name)


# bench lines
for src, dst in (("bench", "r02_bench_c2.json"), ("bench_ref", "r02_bench_c2_reference_arm.json"), ("bench_c1", "r02_bench_c1.json"),
                 ("bench_c2cn", "r02_bench_c2cn.json"), ("bench_mnist_img", "r02_bench_mnist_img.json"),
                 ("bench_c5_sweep", "r02_bench_c5_sweep_full.json")):
    p = os.path.join(G, f"{T}_{src}.json")
    if os.path.exists(p) and os.path.getsize(p):
        shutil.copy(p, os.path.join(P, dst))
        print("copied", dst)
for src, dst in (("small_batch.log", "r02_small_batch.log"), ("train_breakdown.log", "r02_train_breakdown.log"),
                 ("conv_probe.log", "r02_conv_pix_probe.log")):
    p = os.path.join(G, f"{T}_{src}")
    if os.path.exists(p):
        shutil.copy(p, os.path.join(P, dst))

# launch lists
for src, dst, title in (("launches_c2.csv", "r02_launches_c2.md", "C2 log_prob: the two timed steps (cudaProfilerStart/Stop range), fp32 mode, through the whole-stack C entry"),
                        ("launches_train.csv", "r02_launches_train.md", "C3 training step (8192 rows, hand-written pass + SophiaG), one step launch by launch"),
                        ("launches_mnist_img.csv", "r02_launches_mnist_img.md", "image-shaped flow, the reference's live MNIST configuration ([16,7,7], B = 15, ConvNet2D 32 ch x 3 gated 3x3 blocks), ONE log_prob step over 16 384 images (two chunks), the conditioners on pixel planes (usf_conv2d_pix)")):
    p = os.path.join(G, f"{T}_{src}")
    if os.path.exists(p):
        body = run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "launches", p])
        write(dst, f"# r02 — ncu launch list: {title}\n\n`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` "
                   f"(per-launch times are cold-cache and serialised: compare SHARES with the CUDA-event breakdown of the bench line)\n\n" + body)

# full captures
def raw_rows(path):
    """rows / column index / unit row of a `ncu -i <rep> --page raw --csv` dump"""
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return [], {}, []
    hdr = rows[0]
    return rows[2:], {h: i for i, h in enumerate(hdr)}, rows[1]


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__cluster_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]


for src, dst, title in (("full_gemm_raw.csv", "r02_full_gemm_c2_fp32.md", "the 21 contraction launches of ONE C2 log_prob step, fp32 mode (fp16-split x3)"),
                        ("full_gemm_bf16_raw.csv", "r02_full_gemm_c2_bf16.md", "the 21 contraction launches of ONE C2 log_prob step, bf16 mode (TMA store path for bf16 / fp32 planes)"),
                        ("full_train_raw.csv", "r02_full_train_kernels.md", "glue / weight-side kernels of the training step"),
                        ("pix_plain_raw.csv", "r02_full_conv_pix_plain.md", "usf_conv2d_pix, plain 3x3 convolution 32 -> 32 channels over 802 816 pixels (fp32 rows + pixel planes out), tools/pix_ncu.py"),
                        ("pix_gated_raw.csv", "r02_full_conv_pix_gated.md", "usf_conv2d_pix, GatedConv block (3x3 conv + 1x1 conv + gate + ReLU + LayerNorm) over 802 816 pixels, tools/pix_ncu.py")):
    path = os.path.join(G, f"{T}_{src}")
    if not os.path.exists(path):
        continue
    rows, idx, units = raw_rows(path)
    if not rows:
        continue

    def val(r, k):
        try:
            return float(r[idx[k]].replace(",", ""))
        except (ValueError, KeyError):
            return 0.0
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ur, uw = mult.get(units[idx["dram__bytes_read.sum"]], 1), mult.get(units[idx["dram__bytes_write.sum"]], 1)
    tot = [val(r, "dram__bytes_read.sum") * ur + val(r, "dram__bytes_write.sum") * uw for r in rows]
    dur_u = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0)
    dur = [val(r, "gpu__time_duration.sum") * dur_u for r in rows]
    tens = [val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") for r in rows]
    lines = [f"# r02 — ncu --set full: {title}", "",
             f"Sum over the {len(rows)} captured launches: DRAM traffic {sum(tot) / 1e9:.3f} GB, duration {sum(dur) / 1e3:.3f} ms, "
             f"tensor-pipe active (duration-weighted) {sum(t * d for t, d in zip(tens, dur)) / max(sum(dur), 1e-9):.1f}%.", "",
             "| # | kernel | " + " | ".join(k.split(".")[0].replace("__", " ").strip() for k in KEYS) + " |",
             "|---|---|" + "---:|" * len(KEYS)]
    import re
    for n, r in enumerate(rows):
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("usf::", "")
        lines.append(f"| {n} | `{name}` | " + " | ".join(r[idx[k]] + " " + units[idx[k]] if k in idx else "" for k in KEYS) + " |")
    write(dst, "\n".join(lines) + "\n")
    if src == "full_gemm_raw.csv":
        with open(os.path.join(P, "r02_traffic.json"), "w") as f:
            json.dump(dict(workload="c2", precision="fp32", launches=len(rows), dram_bytes_per_launch=sum(tot) / len(rows),
                           dram_bytes_per_step=sum(tot),
                           source="profiles/r02_full_gemm_c2_fp32.md (ncu --set full, the 21 tc2::gemm_tc2_kernel launches of one C2 "
                                  "log_prob step inside a cudaProfilerStart/Stop range)",
                           algorithmic_bytes_per_step=65536 * (4 * 784 + 4)), f, indent=1)
        print("wrote profiles/r02_traffic.json")
