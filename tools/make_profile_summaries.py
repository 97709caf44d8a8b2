#!/usr/bin/env python
"""Turns the ncu launch lists of tools/gpu.sh launches:N (gpurun_out/launches_b1.csv, launches_b4.csv) into the tracked summaries under profiles/:
  profiles/launches_<tag>_b<N>_summary.md   per-launch table of one forward step (time, tensor-pipe %, DRAM MB) + conv-trunk aggregate
  profiles/roofline_traffic_<tag>.json      DRAM bytes of the fc6 launch per batch size (read by bench.py for roofline.traffic)
usage: python tools/make_profile_summaries.py r08"""
import collections, csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONV_NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3"]


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    L = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = L.setdefault(r["ID"], {"name": re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"])).strip(), "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"].startswith("gpu__time"):
            v = v / 1000.0 if u.startswith("n") else (v * 1000.0 if u.startswith("m") else v)
        d[r["Metric Name"]] = v
    ids = list(L)
    starts = [i for i, k in enumerate(ids) if "conv_first" in L[k]["name"]]
    return [L[k] for k in ids[starts[-1]:]]


def main():
    tag = sys.argv[1]
    traffic = {"fc6": {}}
    for B in (1, 4):
        path = os.path.join(ROOT, "gpurun_out", "launches_b%d.csv" % B)
        if not os.path.exists(path):
            continue
        step = load(path)
        T, TP, RD, WR = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"
        total = sum(d[T] for d in step)
        out = ["# Launch list of one forward step, batch %d (%s)\n" % (B, tag),
               "Command (B200 box): `ncu --metrics %s,%s,%s,%s --clock-control none --csv python bench.py --steps 1 --warmup 3 --batch %d "
               "--no-cpu-baseline --no-extras` (tools/gpu.sh launches:%d; raw CSV in gpurun_out/, not tracked).  720x600 images, 1000 proposals, 15 tokens.  "
               "Per-launch times under ncu are cold-cache and serialised (side-stream overlap and programmatic dependent launch are lost): compare SHARES with "
               "the live CUDA-event phase times of bench.py.\n" % (T, TP, RD, WR, B, B),
               "%d launches, %.1f us serialised under ncu.\n" % (len(step), total)]
        # conv trunk aggregate: a layer = conv_first / a halo launch, plus the launches that finish it (split-K finalize; the K-split TAIL launch of the
        # leftover tiles, template argument TAIL = 1, and its conv_tail_finalize).  The 14th layer is the RPN 3x3 conv, not part of the VGG trunk.
        layers = []
        for d in step:
            nm = d["name"]
            is_tail = "tc_conv_halo" in nm and nm.rstrip(">").endswith(", 1")
            if "conv_first" in nm or ("tc_conv_halo" in nm and not is_tail):
                layers.append([d])
            elif layers and (is_tail or "splitk_finalize" in nm or "conv_tail_finalize" in nm):
                layers[-1].append(d)
            elif layers:
                break
        trunk = layers[:13]
        all_t = sum(d[T] for L in trunk for d in L)
        tw = sum(d[T] * d.get(TP, 0.0) for L in trunk for d in L) / max(all_t, 1e-9)
        out.append("## VGG-16 conv trunk (13 layers)\n")
        out.append("| layer | kernels | us | tensor-pipe active % (time-weighted over the layer's launches) |")
        out.append("|---|---|---:|---:|")
        for n, L in zip(CONV_NAMES, trunk):
            t = sum(d[T] for d in L)
            tp = sum(d[T] * d.get(TP, 0.0) for d in L) / max(t, 1e-9)
            out.append("| %s | %s | %s | %.1f |" % (n, " + ".join("`%s`" % d["name"] for d in L), " + ".join("%.1f" % d[T] for d in L), tp))
        out.append("\nTime-weighted tensor-pipe activity over all conv-trunk launches (13 layers incl. their tail / finalize launches): **%.1f %%** "
                   "(%.1f us of kernels; the tensor pipe executes 3 fp16 MMAs per fp32-class product).\n" % (tw, all_t))
        out.append("## All launches\n")
        out.append("| # | kernel | grid | us | share %% | tensor-pipe %% | DRAM read MB | DRAM write MB |".replace("%%", "%"))
        out.append("|---:|---|---|---:|---:|---:|---:|---:|")
        for i, d in enumerate(step):
            out.append("| %d | `%s` | %s | %.1f | %.1f | %.1f | %.1f | %.1f |" % (i, d["name"], d["grid"].strip("()"), d[T], 100.0 * d[T] / total, d.get(TP, 0.0),
                                                                                 d.get(RD, 0.0) / 1e6, d.get(WR, 0.0) / 1e6))
        open(os.path.join(ROOT, "profiles", "launches_%s_b%d_summary.md" % (tag, B)), "w").write("\n".join(out) + "\n")
        # fc6 = the longest tc_gemm launch
        g = [i for i, d in enumerate(step) if "tc_gemm_kernel" in d["name"]]
        if g:
            i0 = max(g, key=lambda i: step[i][T])
            parts = [step[i0]]
            for d in step[i0 + 1:]:        # the K-split tail launch of the leftover tiles and its finalize kernel belong to the same layer
                if ("tc_gemm_kernel" in d["name"] and d["name"].rstrip(">").endswith(", 1")) or "gemm_tail_finalize" in d["name"]:
                    parts.append(d)
                else:
                    break
            tt = sum(d[T] for d in parts)
            traffic["fc6"]["batch%d" % B] = {"kernels": [d["name"] for d in parts], "dram_bytes_read": sum(d.get(RD, 0.0) for d in parts),
                                             "dram_bytes_write": sum(d.get(WR, 0.0) for d in parts), "duration_us_under_ncu": tt,
                                             "tensor_pipe_active_pct": sum(d[T] * d.get(TP, 0.0) for d in parts) / max(tt, 1e-9),
                                             "source": "profiles/launches_%s_b%d_summary.md" % (tag, B)}
        traffic.setdefault("conv_trunk", {})["batch%d" % B] = {"tensor_pipe_active_pct_time_weighted": round(tw, 2), "kernel_us_under_ncu": round(all_t, 1),
                                                               "launches": sum(len(L) for L in trunk), "source": "profiles/launches_%s_b%d_summary.md" % (tag, B)}
        print("batch %d: %d launches, trunk tensor-pipe %.1f %% time-weighted" % (B, len(step), tw))
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "roofline_traffic_%s.json" % tag), "w"), indent=1)


if __name__ == "__main__":
    main()
