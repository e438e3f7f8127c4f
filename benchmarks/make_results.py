"""Render RESULTS.md from the bench JSON lines kept under profiles/ (r01_bench_n{1,2,4,8}.json, r01_bench_reference.json)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
P = ROOT / "profiles"


def load(name):
    f = P / name
    if not f.exists():
        return None
    return json.loads(f.read_text().strip().splitlines()[-1])


def main():
    n1 = load("r01_bench_n1.json")
    ref = load("r01_bench_reference.json")
    out = ["# RESULTS — measured on B200 (round 1)", "",
           "All numbers are CUDA-event timings from `bench.py` / `benchmarks/extras.py` on the pool's B200 boxes "
           "(SM clock %s MHz of %s, throttle reasons: %s). Roofline fractions use the measured copy bandwidth "
           "**%.0f GB/s** (`MEASURED_PEAKS.json`); the nominal figure in BASELINE.json is ~8 TB/s. "
           "Raw JSON lines: `profiles/r01_bench_*.json`; ncu evidence: `profiles/r01_prof_*`." % (
               n1["clocks"]["sm_mhz"], n1["clocks"]["sm_max_mhz"], n1["clocks"]["reasons"] or "none", n1["roofline"]["peak"]), ""]
    out += ["## Headline: BASELINE config 2 (f32 512^3 Zip over non-contiguous views)", "",
            "| | value |", "|---|---|",
            f"| device-resident throughput (`value`) | **{n1['value']:.0f} GB/s** algorithmic = {n1['value'] / n1['roofline']['peak']:.3f} of measured peak, {n1['ms_per_step']:.3f} ms per step (3 Zips) |"]
    for nm, d in n1["roofline"]["per_case"].items():
        out.append(f"| &nbsp;&nbsp;{nm} ({d['kernel_class']}) | {d['GB/s']:.0f} GB/s ({d['GB/s'] / n1['roofline']['peak']:.3f}), {d['ms'] * 1e3:.1f} µs |")
    out += [f"| dominant kernel roofline | {n1['roofline']['kernel']}: {n1['roofline']['achieved']:.0f} / {n1['roofline']['peak']:.0f} GB/s = **{n1['roofline']['frac']:.3f}**; ncu DRAM traffic {n1['roofline']['traffic']} B vs 1 610 612 736 algorithmic |",
            f"| end to end through `ndb_zip_host` (pinned host buffers, H2D + kernel + D2H) | **{n1['e2e']['value']:.1f} GB/s** algorithmic; {n1['e2e']['h2d_bytes_per_step'] / 1e9:.2f} GB up + {n1['e2e']['d2h_bytes_per_step'] / 1e9:.2f} GB down per step in {n1['e2e']['ms_per_step']:.1f} ms — the box's PCIe ceiling is 55.6 GB/s one way, 45 GB/s each way when both directions run (`benchmarks/pcie_probe.py`) |"]
    if ref:
        out.append(f"| reference arm (`--impl reference`: oracle port of ndarray's traversal, {ref['cpu_baseline']['cores']} host threads, full 512^3) | {ref['value']:.1f} GB/s ({ref['ms_per_step']:.0f} ms per step) |")
    out.append(f"| cpu_baseline sample inside the own-arm run | {n1['cpu_baseline']['value']:.1f} GB/s — {n1['cpu_baseline']['sample']} |")
    out += ["", "## Scaling (weak: one 512^3 problem per GPU, no data-path collective)", "", "| GPUs | aggregate GB/s | × 1 GPU | e2e GB/s |", "|---|---|---|---|"]
    for n in (1, 2, 4, 8):
        d = load(f"r01_bench_n{n}.json")
        if d:
            out.append(f"| {n} | {d['value']:.0f} | {d['value'] / n1['value']:.2f} | {d['e2e']['value']:.0f} |")
    d8 = load("r01_bench_n8_extras.json") or load("r01_bench_n8.json")  # the run WITH extras (headline lines above are --no-extras reruns)
    if d8 and "extra" in d8 and "C5_zip+global_sum_f32" in d8["extra"]:
        ex = d8["extra"]
        c5 = ex["C5_zip+global_sum_f32"]
        out += ["", "## Multi-GPU exchange steps (8 GPUs, `extra` of `profiles/r01_bench_n8_extras.json`)", "",
                "| case | NCCL (`torch.distributed` on the ctx stream) | peer memory over NVLink (`csrc/peer.cu`) |", "|---|---|---|"]
        if "C5_zip+global_sum_f32_peer_fused" in ex:
            pf = ex["C5_zip+global_sum_f32_peer_fused"]
            out.append(f"| config 5 shape: per rank 2^29 f32, `a = b*c + a` then global `sum()` | {c5['GB/s_aggregate']:.0f} GB/s aggregate, {c5['ms']} ms | "
                       f"**{pf['GB/s_aggregate']:.0f} GB/s aggregate**, {pf['ms']} ms (exchange fused into the sum's finishing kernel) |")
        for lab, name in (("1elt", "all-reduce of one element"), ("128KiB", "all-reduce of 128 KiB (config 3's `sum_axis(Axis(0))` vector)")):
            k = f"allreduce_latency_{lab}"
            if k in ex:
                out.append(f"| {name} | {ex[k]['nccl_us']} µs | **{ex[k]['peer_memory_us']} µs** |")
        if "C3_sum_axis0_sharded_f64_nccl" in ex:
            a, b = ex["C3_sum_axis0_sharded_f64_nccl"], ex["C3_sum_axis0_sharded_f64_peer"]
            f = ex.get("C3_sum_axis0_sharded_f64_peer_fused_one_launch")
            out.append(f"| sharded `sum_axis(Axis(0))`, 2048x16384 f64 per rank + exchange | {a['GB/s_aggregate']:.0f} GB/s aggregate ({a['ms'] * 1e3:.0f} µs) | "
                       f"**{b['GB/s_aggregate']:.0f} GB/s aggregate** ({b['ms'] * 1e3:.0f} µs) with reduce + exchange kernel"
                       + (f"; {f['GB/s_aggregate']:.0f} GB/s ({f['ms'] * 1e3:.0f} µs) as ONE fused launch (`ndb_reduce_axis_peer`)" if f else "") + " |")
        if "C3_sum_axis0_sharded_f64_var_nccl" in ex:
            a, b = ex["C3_sum_axis0_sharded_f64_var_nccl"], ex["C3_sum_axis0_sharded_f64_var_peer_fused_one_launch"]
            out.append(f"| sharded `var_axis(Axis(0), 1.0)`, same slab: Welford states of all ranks merged in rank order | {a['GB/s_aggregate']:.0f} GB/s aggregate "
                       f"({a['ms'] * 1e3:.0f} µs: local kernel, three all-gathers, merge kernel) | **{b['GB/s_aggregate']:.0f} GB/s aggregate** "
                       f"({b['ms'] * 1e3:.0f} µs, ONE launch per rank) |")
        out += ["", "Row-sharded `general_mat_mul` (A and C sharded along Axis(0), 8192 rows per rank; B broadcast from rank 0 with "
                    "`ncclBroadcast` INSIDE every timed step):", "", "| case | aggregate | per step |", "|---|---|---|"]
        for k, v in ex.items():
            if "rowsharded" in k:
                out.append(f"| {k} | {v['TFLOP/s_aggregate']} TFLOP/s | {v['ms']} ms |")
    out += ["", "## Other BASELINE configs (1 GPU, `extra` in the bench line)", "", "| case | measured | of peak |", "|---|---|---|"]
    for k, v in n1.get("extra", {}).items():
        if "GB/s" in v:
            out.append(f"| {k} | {v['GB/s']} GB/s ({v['ms'] * 1e3:.1f} µs) | {v['frac_of_measured_peak']} |")
        elif "TFLOP/s" in v:
            out.append(f"| {k} | {v['TFLOP/s']} TFLOP/s" + (f" ({v['ms']} ms)" if "ms" in v else "") + " | |")
    out += ["", "GEMM context: the TMA-fed f64 kernel reaches %.2f x cuBLAS `Dgemm` on the same box (37.3 TFLOP/s is the FP64 tensor pipe's peak at 1965 MHz); the 3xTF32 kernel issues three tensor "
            "MMAs per logical MMA, so its tensor-pipe rate is 3× its reported figure — compare with the cuBLAS 1xTF32 line: the pipe "
            "is power-limited at about that rate (`NDB_GEMM_DEBUG=1`, which removes all TMA traffic, does not speed the kernel up)." % (
                n1["extra"]["C4_gemm_f64_8192^3_CCC"]["TFLOP/s"] / n1["extra"]["C4_comparator_cublas_f64_8192^3"]["TFLOP/s"]) if "extra" in n1 and "C4_gemm_f64_8192^3_CCC" in n1["extra"] else ""]
    (ROOT / "RESULTS.md").write_text("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
