"""profiles/roofline.json from one `ncu --set full` capture of step_kernel<true>:  python tools/update_roofline.py <rep> <n_envs> <tag>
fp32 work = thread-level executed FFMA (x2) + FADD + FMUL (predicated-on); traffic = dram read + write bytes of that launch."""
import csv, json, os, subprocess, sys
rep, n, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
m = dict(zip(rows[0], rows[-1])); u = dict(zip(rows[0], rows[1]))
f = lambda k: float(m[k].replace(",", ""))
cyc = f("smsp__cycles_elapsed.avg")
ffma, fadd, fmul = (f(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") * cyc for k in ("ffma", "fadd", "fmul"))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd = f("dram__bytes_read.sum") * scale[u["dram__bytes_read.sum"]]; wr = f("dram__bytes_write.sum") * scale[u["dram__bytes_write.sum"]]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "roofline.json")
d = json.load(open(path))
dram = dict(d.get("dram_bytes_per_launch", {})); dram[str(n)] = round(rd + wr)  # one entry per profiled env count
wi = dict(d.get("warp_instructions_per_env_step", {})); wi[str(n)] = round(f("smsp__inst_executed.sum") / n, 1)
d.update({"fp32_flops_executed_per_env_step": round((2 * ffma + fadd + fmul) / n), "profile": tag, "profile_n_envs": n,
          "dram_bytes_per_launch": dram, "warp_instructions_per_env_step": wi, "dram_read_write_MB": [round(rd / 1e6, 1), round(wr / 1e6, 1)],
          "kernel_time_under_ncu_us": f("gpu__time_duration.sum") * (1e3 if u["gpu__time_duration.sum"] == "ms" else 1.0),
          "how_executed": "thread-level executed FFMA (x2) + FADD + FMUL of step_kernel<true> in the profiled launch (ncu --set full), per env. Since the "
                          "Newton loop became warp-synchronous (r1h) this count includes the masked trips of lanes whose solve has finished, so it is "
                          "NOT the algorithmic figure; fp32_flops_per_env_step stays frozen at the r1c value, measured when idle lanes were predicated off.",
          "dram": "dram__bytes_read.sum + dram__bytes_write.sum of the same launch"})
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d, indent=1))
