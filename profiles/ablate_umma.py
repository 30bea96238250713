import os, sys, subprocess
for ab in (0, 1, 2, 4, 8, 3, 6, 7, 15):
    env = dict(os.environ, CA_OSC_ABLATE=str(ab))
    out = subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-k", "regex:osc_umma_kernel", "-s", "3", "-c", "1", "--csv",
                          sys.executable, "profiles/prof_driver.py", "osc"], env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if "osc_umma_kernel" in l and "gpu__time" in l]
    print("ablate", ab, line[0].split(",")[-1] if line else out[-300:], flush=True)
