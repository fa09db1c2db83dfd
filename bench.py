#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused CyGym step kernel on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the CPU arm (oracle port on the host cores)

Workload (config.workload "C3"): 65 536 envs x 100 device slots / 8 subnets per GPU, synthetic network
(cygym_b200.network.synthetic_network), uniformly random sample_action()-style defender / attacker
actions on alternating turns, detector untrained (defender action 10 rewritten to the no-op 8).
A "step" is ONE launch of cyg_step over one batch of B envs.  Envs shard across GPUs with no
data-path collective (weak scaling: B envs per GPU); NCCL only carries the timing reduction.

L2 policy: the bench rotates over `--sets` independent env sets (default 3 x 57 MB of state
> 126 MB L2) and a ring of pre-generated action batches, so a launch never finds its records
in L2 from the previous launch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec at 64K envs x 100 devices"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--devices", type=int, default=100)
    ap.add_argument("--subnets", type=int, default=8)
    ap.add_argument("--sets", type=int, default=3, help="independent env sets rotated to defeat L2 residency")
    ap.add_argument("--ring", type=int, default=8, help="pre-generated action batches per mode")
    ap.add_argument("--fuse", type=int, default=4, help="plain steps fused per launch (VectorCyberDefenseEnv.step_many / cyg_step_multi); 1 = one launch per step")
    ap.add_argument("--randomize", action="store_true", help="randomize_compromise_and_ownership() on every env set first (diagnostic: "
                    "the owned set moves, evolve_network adds extra hub-star edges, and the steps take the kernels' extra-edge forms)")
    ap.add_argument("--obs", type=int, default=0, help="fused observation mode inside the step (0 none, 1 defender, 2 attacker)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


def workload_name(a):
    return f"C3: {a.envs} envs x {a.devices} device slots / {a.subnets} subnets per GPU, random sample_action defender/attacker turns"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region.  The timed region is tens of milliseconds, so the
    sampler polls NVML in-process every ~2 ms (nvidia-smi, the B200_PROFILING.md clocks line, is the fallback)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.max_sm, self.reasons, self.how = [], 0.0, set(), "nvml"
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[gpu_index])
                                                        if os.environ.get("CUDA_VISIBLE_DEVICES") else gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            self.how = "nvidia-smi"

    def _poll_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _poll_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            s = [x.strip() for x in out.split(",")]
            self.sm.append(float(s[1]))
            self.max_sm = max(self.max_sm, float(s[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._poll_nvml() if self._nvml else self._poll_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.002 if self._nvml else 0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm or None, "reasons": sorted(self.reasons),
                "samples": len(sm), "how": self.how}


def cpu_arm(a, seconds, threads=None):
    """The oracle port (oracle/cyg_oracle.c, OpenMP over envs) timed on the host cores on a bounded
    sample of the same workload.  Returns (env-steps/s, cores, description)."""
    import numpy as np
    from cygym_b200 import synthetic_network
    from oracle import cyg_oracle as O
    from tests.common import oracle_for, oracle_state_from_template, sanitize_actions
    net = synthetic_network(a.devices, n_subnets=a.subnets, seed=a.seed)
    cores = threads or os.cpu_count() or 1
    cores = min(cores, O.lib().cyo_max_threads()) if O.lib().cyo_max_threads() > 0 else 1
    Bc = max(cores * 256, 1024)
    orc, _ = oracle_for(net, seed=a.seed, xcap=16)
    st = oracle_state_from_template(orc, net, Bc)
    acts = []
    for mode in (0, 1):
        h, m = orc.sample_actions(st, mode)
        h = sanitize_actions(h, np.ones(Bc, np.uint32), mode)
        acts.append((h, m))
    for t in range(4):
        orc.step(st, *acts[t & 1], n_threads=cores)
    n, t0 = 0, time.perf_counter()
    while True:
        orc.step(st, *acts[n & 1], n_threads=cores)
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 100000:
            break
    return Bc * n / el, cores, f"{Bc} envs x {n} steps of the same workload, {cores} OpenMP threads, {el:.1f}s"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step = max(0.5, min(20.0, 120.0 / max(1, a.steps + a.warmup)))
    # each "step" of this arm is one bounded CPU sample
    vals = []
    for i in range(a.warmup + a.steps):
        v, cores, sample = cpu_arm(a, per_step)
        if i >= a.warmup:
            vals.append(v)
    val = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "l2_policy": "n/a (CPU arm)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + f" per step, {a.steps} steps"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    import numpy as np
    import torch
    import torch.distributed as dist
    from cygym_b200 import synthetic_network
    from cygym_b200.vector_env import ActionBatch, VectorCyberDefenseEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = a.envs, a.steps, max(3, a.warmup)

    net = synthetic_network(a.devices, n_subnets=a.subnets, seed=a.seed)
    # one env set per rotation slot; env ids are globally unique across ranks and sets
    sets = [VectorCyberDefenseEnv(net, B, device=dev, seed=a.seed, env_id0=(rank * a.sets + s) * B, xcap=16)
            for s in range(a.sets)]
    if a.randomize:
        for s_ in sets:
            s_.randomize_compromise_and_ownership()
    # ring of pre-generated action batches (inputs resident in HBM before the timed region)
    ring = {0: [], 1: []}
    for mode in (0, 1):
        for r in range(a.ring):
            ab = sets[r % a.sets].sample_actions(mode)
            if mode == 0:  # detector untrained: defender 10 -> no-op 8
                ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
            for spec in filter(None, os.environ.get('CYG_BENCH_EXCLUDE', '').split(',')):  # diagnostics only
                m_, t_ = spec.split(':')
                if int(m_) == mode:
                    ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == int(t_), (ab.hdr[:, 0] & ~0xFF) | (8 if mode == 0 else 3), ab.hdr[:, 0])
            ring[mode].append(ActionBatch(ab.hdr.clone(), ab.mask.clone()))
    torch.cuda.synchronize()

    def one_step(i):
        env = sets[i % a.sets]
        turn = (i // a.sets) & 1  # every set alternates defender / attacker turns
        env.step(ring[turn][(i // (2 * a.sets)) % a.ring], obs_mode=a.obs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, warm):
        for i in range(warm):
            fn(i)
        barrier()
        l0 = sum(s.launch_count for s in sets)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(n):
            fn(warm + i)
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1), sum(s.launch_count for s in sets) - l0

    F = max(1, a.fuse) if not a.obs else 1
    sampler = ClockSampler(local)
    if F > 1:
        # the headline leg: F plain steps per launch (alternating defender / attacker turns), the records of a CTA's envs
        # stay in shared memory between them; every env set gets its own [F, B, ..] action tensors out of the ring
        K = ((K + F - 1) // F) * F
        fused = []
        for s_i in range(a.sets):
            per_set = []
            for v in range(2):  # two variants per set so that consecutive launches of a set read different batches
                hs = [ring[t & 1][(s_i + v * 3 + t // 2) % a.ring].hdr for t in range(F)]
                mk = [ring[t & 1][(s_i + v * 3 + t // 2) % a.ring].mask for t in range(F)]
                per_set.append((torch.stack(hs).contiguous(), torch.stack(mk).contiguous()))
            fused.append(per_set)
        outs = [(torch.empty(F, B, dtype=torch.float32, device=dev), torch.empty(F, B, dtype=torch.float32, device=dev),
                 torch.empty(F, B, dtype=torch.int32, device=dev)) for _ in range(a.sets)]
        torch.cuda.synchronize()

        def one_launch(j):
            s_i = j % a.sets
            h_, m_ = fused[s_i][(j // a.sets) & 1]
            sets[s_i].step_many(h_, m_, out=outs[s_i])

        sampler.start()
        ms, launches = timed(one_launch, K // F, max(1, (Wm + F - 1) // F))
        clocks = sampler.stop()
        K1 = min(K, 600)
        ms1, _ = timed(one_step, K1, Wm)  # the same workload, one launch per step
    else:
        sampler.start()
        ms, launches = timed(one_step, K, Wm)
        clocks = sampler.stop()
        ms1, K1 = ms, K
    errs = int(max(int(s.error_flags().max().item()) for s in sets))

    # ---- e2e: the public host-buffer call VectorCyberDefenseEnv.step_host(): every step copies that step's actions
    #      from pinned host memory, launches the kernel, reads (raw, shaped, done) back and synchronises ----
    e2e = None
    if not a.no_e2e:
        Ke = max(10, min(K, 200))
        groups = sets[:3]  # every env set of the rotation is one group with its own stream
        host_actions = {}
        for gi, env in enumerate(groups):
            env._stream = torch.cuda.Stream(dev)  # one stream per env group
            env.host_buffers()
            for mode in (0, 1):
                ab = ring[mode][gi % len(ring[mode])]
                host_actions[gi, mode] = torch.cat([ab.hdr, ab.mask], dim=1).cpu().pin_memory()  # [B, 4 + W] rows: hdr | mask
        out_host = groups[0].host_buffers()[2]
        torch.cuda.synchronize()

        def run_e2e(n_groups, steps):
            """steps host-buffer steps over n_groups env groups (round robin).  Every step: wait for that group's
            previous results (the caller reads them before it chooses the group's next action), then enqueue H2D of the
            step's actions, the kernel and the D2H of (raw, shaped, done)."""
            turn = [0] * n_groups
            for g in range(n_groups):
                groups[g].wait_host()
            barrier()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(steps):
                g = i % n_groups
                env = groups[g]
                env.wait_host()
                env.step_host(act=host_actions[g, turn[g] & 1], sync=False)
                turn[g] += 1
            for g in range(n_groups):
                groups[g].wait_host()
            t1.record()
            barrier()
            return t0.elapsed_time(t1)

        run_e2e(len(groups), 6)
        ems = run_e2e(len(groups), Ke)
        run_e2e(1, 4)
        ems1 = run_e2e(1, max(10, Ke // 2))
        e2e = (ems, Ke, host_actions[0, 0].numel() * 4, out_host.numel() * 4, ems1, max(10, Ke // 2), len(groups))

    # ---- max over ranks ----
    t = torch.tensor([ms, e2e[0] if e2e else 0.0, e2e[4] if e2e else 0.0, ms1], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ems, ems1, ms1 = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg = net.algorithmic_bytes_per_step(obs=bool(a.obs))
        traffic = None  # dram read+write bytes per launch of the step kernel from the committed `ncu --set full` capture
        try:
            with open(os.path.join(ROOT, "profiles", "r01_step_kernel_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("envs_per_launch") == B and tj.get("device_slots") == a.devices and not a.obs:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        traffic_f = None  # the same for a fused launch (its own capture)
        try:
            with open(os.path.join(ROOT, "profiles", "r01_step_kernel_fused_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("envs_per_launch") == B and tj.get("device_slots") == a.devices and tj.get("steps_per_launch") == F:
                traffic_f = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        launch_s = ms * 1e-3 / (K // F)      # average duration of one launch (F steps)
        achieved = alg * B * F / launch_s / 1e9
        value = world * B * K / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(a), "envs_per_gpu": B, "device_slots": a.devices, "edges": net.E,
                       "obs_mode": a.obs, "env_sets": a.sets, "record_bytes": sets[0].S * 4, "actions": "pre-generated ring of sample_action batches resident in HBM; defender 10 (detector training) -> no-op",
                       "randomized_ownership": bool(a.randomize), "steps_per_launch": F,
                       "fusion": (f"{F} plain steps per launch (step_many / cyg_step_multi): open-loop action batches resident in HBM, records stay in "
                                  "shared memory between the steps of a launch, so per-step HBM traffic is actions in + rewards out; "
                                  "single_step_launch below is the same workload at one launch per step") if F > 1 else "one launch per step",
                       "l2_policy": f"rotating {a.sets} env sets ({a.sets * B * (sets[0].S + net.M) * 4 / 1e6:.0f} MB of state > 126 MB L2) and {2 * a.ring} action batches",
                       "parallelism": f"env-sharded x{world}, no per-step collective", "error_flags": errs},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic if F == 1 else traffic_f, "algorithmic_bytes_per_env_step": alg, "envs_per_launch": B,
                         "env_steps_per_launch": B * F, "launch_us": launch_s * 1e6, "peak_source": peak_src,
                         "actual_bytes_per_env_step": (2 * sets[0].S * 4) / F + 4 * (4 + net.W) + 12,
                         "note": ("algorithmic bytes are SURVEY 8(d)'s per-step figure x env-steps per launch; a fused launch moves less than that "
                                  "(state in/out once per launch), which 8(d) allows for") if F > 1 else None},
            "single_step_launch": {"value": world * B * K1 / (ms1 * 1e-3), "unit": UNIT, "launch_us": ms1 * 1e3 / K1, "steps": K1,
                                   "roofline_frac": alg * B / (ms1 * 1e-3 / K1) / 1e9 / peak, "traffic": traffic},
            "clocks": clocks, "gpu_launches": int(launches),
        }
        if e2e:
            line["e2e"] = {"value": world * B * e2e[1] / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e[2],
                           "d2h_bytes_per_step": e2e[3], "steps": e2e[1], "ms_per_step": ems / e2e[1],
                           "how": f"VectorCyberDefenseEnv.step_host(act=pinned rows, sync=False) / wait_host() over {e2e[6]} env groups of "
                                  f"{B} envs on {e2e[6]} streams: each group waits for its own previous (raw, shaped, done) before its next "
                                  "step; the copies of one group overlap the kernel of the other",
                           "one_group_synchronous": {"value": world * B * e2e[5] / (ems1 * 1e-3), "unit": UNIT,
                                                     "ms_per_step": ems1 / e2e[5], "steps": e2e[5]}}
        if not a.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_arm(a, a.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
