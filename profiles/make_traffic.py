"""Rebuild profiles/traffic.json (the constants bench.py copies into roofline.traffic / roofline.ncu / extra.synthesis.*.ncu) from the
committed ncu summaries, so that the two cannot drift apart:   python profiles/make_traffic.py"""
import json, os
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    return json.load(open(os.path.join(HERE, name)))


def num(s):
    v, _, unit = s.partition(" ")
    return float(v), unit.strip()


def bytes_of(s):
    v, unit = num(s)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def ms_of(s):
    v, unit = num(s)
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]


def counters(d, keys):
    names = {"fma_pipe_cycles_active_pct": "fma_pipe_cycles_active_pct", "fma_pipe_pct": "fma_pipe_inst_pct", "issue_active_pct": "issue_active_pct",
             "l1_wavefronts_pct": "l1_wavefronts_pct", "ipc_per_sm": "ipc_per_sm", "warps_active_pct": "warps_active_pct", "dram_pct2": "dram_pct",
             "tensor_pipe_cycles_active_pct": "tensor_pipe_cycles_active_pct"}
    out = {names[k]: num(d[k])[0] for k in keys if k in d}
    out["duration_ms"] = ms_of(d["duration"])
    return out


stft = load("ncu_stft4096_r02.json")[0]
rein = load("ncu_oscbank_reinsch_r02.json")[0]
umma = {d["kernel"].split("(")[0].split("::")[-1].split("<")[0].strip(): d for d in load("ncu_oscbank_umma_r02.json")}
mma = load("ncu_oscbank_mma_r02.json")[0]
fp32 = load("ncu_oscbank_fp32_r02.json")[0]
u = umma["osc_umma_kernel"]
t = {
    "stft4096p_kernel": int(bytes_of(stft["dram_read"]) + bytes_of(stft["dram_write"])),
    "stft4096p_kernel_counters": dict(source="profiles/ncu_stft4096_r02.json (ncu --set full, one launch over 32 streams x 11235 frames)",
                                      **counters(stft, ["fma_pipe_cycles_active_pct", "fma_pipe_pct", "issue_active_pct", "l1_wavefronts_pct", "ipc_per_sm",
                                                        "warps_active_pct", "dram_pct2"])),
    "oscbank_kernel_counters": dict(source="profiles/ncu_oscbank_reinsch_r02.json (the recurrence kernel this round started from; ncu --set full, config 3)",
                                    **counters(rein, ["fma_pipe_cycles_active_pct", "fma_pipe_pct", "issue_active_pct", "ipc_per_sm", "warps_active_pct"])),
    "_note": "dram__bytes_read.sum + dram__bytes_write.sum for one launch over 32 streams x 11235 frames, and pipe counters of the headline kernels, from "
             "the committed ncu --set full captures (profiles/make_traffic.py); bench.py copies these constants into roofline.traffic / roofline.ncu / "
             "extra.synthesis.*.ncu, they are not measured in the bench run",
    "oscbank_umma_counters": {
        "source": "profiles/ncu_oscbank_umma_r02.json (ncu --set full, config 3: the launches of one render)",
        "osc_umma_kernel": dict(dram_write_bytes=bytes_of(u["dram_write"]),
                                **counters(u, ["dram_pct2", "tensor_pipe_cycles_active_pct", "fma_pipe_cycles_active_pct", "issue_active_pct", "l1_wavefronts_pct",
                                               "ipc_per_sm"])),
        "oscbank_kernel (fill)": {"duration_ms": ms_of(umma["oscbank_kernel"]["duration"])},
        "osc_umma_tables_kernel": {"duration_ms": ms_of(umma["osc_umma_tables_kernel"]["duration"])},
    },
    "oscbank_mma_counters": dict(source="profiles/ncu_oscbank_mma_r02.json (CA_OSC_PATH=mma)",
                                 **counters(mma, ["tensor_pipe_cycles_active_pct", "l1_wavefronts_pct", "issue_active_pct", "ipc_per_sm"])),
    "oscbank_fp32_counters": dict(source="profiles/ncu_oscbank_fp32_r02.json (CA_OSC_PATH=fp32)",
                                  **counters(fp32, ["fma_pipe_cycles_active_pct", "fma_pipe_pct", "l1_wavefronts_pct", "issue_active_pct", "ipc_per_sm"])),
}
json.dump(t, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
print(json.dumps(t, indent=1))
