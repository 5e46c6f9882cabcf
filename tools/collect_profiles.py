"""gpurun_out/ (scratch) -> profiles/ (committed): bench lines, launch list, ncu summaries,
K10 micro-benchmarks, the SASS evidence of the tensor-core kernel.  Run in the build container
after tools/gpu_final.sh came back."""
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def json_lines(path):
    if not os.path.exists(path):
        return []
    return [json.loads(l) for l in open(path) if l.startswith("{")]


def main():
    bench = {}
    for key, f in (("n1_default", "final_bench.log"), ("reference_arm", "final_ref.log"),
                   ("n1_of_scaling_pair", "scale_n1.log"), ("n2_torchrun", "scale_n2.log")):
        lines = json_lines(os.path.join(OUT, f))
        if lines:
            bench[key] = lines[-1]
    ab = {}
    for f in sorted(glob.glob(os.path.join(OUT, "rb_bench_*_*.log"))):
        lines = json_lines(f)
        if lines:
            lin, conv = os.path.basename(f)[len("rb_bench_"):-4].split("_")
            ab["linear=%s,conv=%s" % (lin, conv)] = {k: lines[-1]["rainbow"][k] for k in (
                "env_steps_per_sec", "e2e_env_steps_per_sec", "ms_per_update_incl_acting")}
    if ab:
        bench["rainbow_dense_layer_policy_ab"] = ab
    json.dump(bench, open(os.path.join(PROF, "r02_bench.json"), "w"), indent=1)

    final = os.path.join(OUT, "final_k10.log")
    if os.path.exists(final):  # tools/bench_gemm.py + bench_conv.py + gemm_phases.py of the final run
        recs = json_lines(final)
        k10 = {"gemm": [r for r in recs if "shape" in r], "conv": [r for r in recs if "layer" in r],
               "phases_per_cta": [l.rstrip() for l in open(final) if " CTAs, " in l]}
    else:
        k10 = {"gemm": json_lines(os.path.join(OUT, "gemm_bench.log")),
               "conv": json_lines(os.path.join(OUT, "conv_bench.log"))}
    json.dump(k10, open(os.path.join(PROF, "r02_k10_bench.json"), "w"), indent=1)

    src = os.path.join(OUT, "r02_launches_ncu.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, "r02_launches_ncu.csv"))

    caps = [("r02_step_exact", "step_exact"), ("r02_step_tail", "step_parallel:1"),
            ("r02_step_tail", "step_parallel_u8_out:3"),
            ("r02_sampler_v6", "sampler_v6"), ("r02_gather", "gather"),
            ("r02_update_multi", "update_multi"), ("r02_gemm", "gemm_tf32x3"), ("r02_gae", "gae"),
            ("r02_ppo_loss", "ppo_loss"), ("r02_polyak", "polyak"), ("r02_sac_target", "sac_target"),
            ("r02_conv1_u8", "conv_nature1_u8")]
    args = []
    for rep, name in caps:
        p = os.path.join(OUT, rep + ".ncu-rep")
        if os.path.exists(p):
            args += [p, name]
    if args:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py")] + args,
                             capture_output=True, text=True).stdout
        open(os.path.join(PROF, "r02_ncu_full_summary.csv"), "w").write(
            "capture,metric,value,unit\n" + out)

    # SASS evidence of the tensor-core kernel
    so = os.path.join(ROOT, "pfrl_b200", "csrc", "libb2rl.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    keep, fn = [], None
    counts = {}
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.strip()
        if any(m in line for m in ("UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "SYNCS")):
            if fn and "gemm_tf32x3ILi0ELi0" in fn:
                keep.append(line.rstrip())
            mnem = [m for m in ("UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP") if m in line]
            for m in mnem:
                counts.setdefault(fn, {}).setdefault(m, 0)
                counts[fn][m] += 1
    with open(os.path.join(PROF, "r02_gemm_sass.txt"), "w") as f:
        f.write("# cuobjdump -sass pfrl_b200/csrc/libb2rl.so : tcgen05 / TMEM / bulk-copy mnemonics\n")
        f.write("# per kernel (count of instructions)\n")
        for k, v in counts.items():
            if v:
                f.write("%s  %s\n" % (k, v))
        f.write("\n# k_gemm_tf32x3<K-major, K-major>: the tensor-core / mbarrier instructions in order\n")
        f.write("\n".join(keep) + "\n")
    print("profiles/ refreshed:", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
