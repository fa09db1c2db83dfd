"""Drop-in single-env front end: the reset()/step() Gym surface of the reference's
Volt_Typhoon_CyberDefenseEnv (volt_typhoon_env.py:30) on top of the batched CUDA path (a batch of 1).

Callers in the reference (do_agent.py, IPPO.py, MAPPO.py, HMARL.py, ...) touch the env through the members
listed in SURVEY.md section 8(b); those are what this class provides, with the same names, argument meaning,
return shapes and error behaviour.  The Device / App / ... object graph itself is not rebuilt: state lives
on the GPU as bit-planes and is exposed through the observation vectors and counters the callers read.
"""
import numpy as np
import torch

from . import _capi as K
from .network import Network, synthetic_network
from .vector_env import ActionBatch, VectorCyberDefenseEnv

_COUNTERS = {
    "step_num": K.S_STEP, "defender_step": K.S_DEF_STEP, "attacker_step": K.S_ATT_STEP, "work_done": K.S_WORK,
    "checkpoint_count": K.S_CKPT, "revert_count": K.S_REVERT, "scan_cnt": K.S_SCAN,
    "compromised_devices_cnt": K.S_COMPCNT, "edges_blocked": K.S_EBLK, "edges_added": K.S_EADD,
}
_FLOAT_COUNTERS = {"defensive_cost": K.S_DEFCOST, "clearning_cost": K.S_CLEANCOST}


class Volt_Typhoon_CyberDefenseEnv:
    """reset(from_init) / step(action, agent_cnt) / step_grouped(groups) with the reference's 6-tuple return
    (state, raw_reward, shaped_reward, done, info, logs) (volt_typhoon_env.py:1333)."""

    MaxExploits = 6

    def __init__(self, network: Network = None, device="cuda:0", seed=0, env_id=0, xcap=32):
        self.numOfDevice = 10            # init_experiments.py:41-42 defaults
        self.Max_network_size = 20
        self.mode = "defender"
        self.tech = "DO"
        self.debug = False
        self.zero_day = False
        self.snapshot_path = None
        self.time_budget_deadline = None
        self.time_budget_exceeded = False
        self._device, self._seed, self._env_id, self._xcap = device, seed, env_id, xcap
        self._net = network
        self._venv = None
        self._base_line = "Nash"
        self._scales = {}
        self.state = None
        if network is not None:
            self.numOfDevice, self.Max_network_size = network.cfg["numOfDevice"], network.M
            self._build()

    # ---- configuration attributes the callers set ----
    @property
    def base_line(self):
        return self._base_line

    @base_line.setter
    def base_line(self, name):
        self._base_line = name
        if self._venv is not None:
            self._venv.set_base_line(name)

    def _scale_get(self, k):
        return self._scales.get(k, None if self._net is None else self._net.cfg[k])

    def _scale_set(self, k, v):
        self._scales[k] = float(v)
        if self._venv is not None:  # scales are kernel constants: rebuild the handle, keep the state
            st = self._venv.export_state()
            self._net.cfg[k] = float(v)
            self._venv.close()
            self._venv = VectorCyberDefenseEnv(self._net, 1, device=self._device, seed=self._seed, env_id0=self._env_id,
                                               base_line=self._base_line, xcap=self._xcap)
            self._venv.import_state(st)

    work_scale = property(lambda s: s._scale_get("work_scale"), lambda s, v: s._scale_set("work_scale", v))
    comp_scale = property(lambda s: s._scale_get("comp_scale"), lambda s, v: s._scale_set("comp_scale", v))
    def_scale = property(lambda s: s._scale_get("def_scale"), lambda s, v: s._scale_set("def_scale", v))

    def _build(self):
        for k, v in self._scales.items():
            self._net.cfg[k] = v
        if self._venv is not None:
            self._venv.close()
        self._venv = VectorCyberDefenseEnv(self._net, 1, device=self._device, seed=self._seed, env_id0=self._env_id,
                                           base_line=self._base_line, xcap=self._xcap)
        self.state = self._get_state()

    # ---- setup (volt_typhoon_env.py:1485-1900, :1904-2107) ----
    def initialize_environment(self, seed=None, n_subnets=None):
        """Builds a synthetic network of the configured size (numOfDevice, Max_network_size) and resets to it."""
        M = int(self.Max_network_size)
        self._net = synthetic_network(M, num_of_device=int(self.numOfDevice),
                                      n_subnets=n_subnets or max(1, M // 12), seed=self._seed if seed is None else seed)
        self._build()
        return self.state

    def reset(self, from_init=True):
        if self._venv is None:
            return self.initialize_environment()
        self._venv.reset()
        self.state = self._get_state()
        return self.state

    def seed(self, seed=None):
        self._seed = 0 if seed is None else int(seed)
        return [self._seed]

    # ---- sizes ----
    def get_num_action_types(self, mode=None):
        if mode == "defender":
            return 14
        if mode == "attacker":
            return 3
        raise ValueError("Invalid mode: must be either 'defender' or 'attacker'")

    def get_device_indices(self):
        return list(range(self._net.M))

    def get_num_exploit_indices(self):
        return int(self._net.cfg["n_exploits"])

    def get_num_app_indices(self):
        return int(self._net.cfg.get("n_app_ids", 0))

    # ---- counters: live views of the env's scalars; callers zero them between rollouts (do_agent.py:192-196) ----
    def __getattr__(self, name):
        if name in _COUNTERS and self.__dict__.get("_venv") is not None:
            return int(self._venv.scalars[0, _COUNTERS[name]].item())
        if name in _FLOAT_COUNTERS and self.__dict__.get("_venv") is not None:
            return float(self._venv.scalars[0, _FLOAT_COUNTERS[name]].view(torch.float32).item())
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in _COUNTERS and self.__dict__.get("_venv") is not None:
            self._venv.scalars[0, _COUNTERS[name]] = int(value)
        elif name in _FLOAT_COUNTERS and self.__dict__.get("_venv") is not None:
            self._venv.scalars[0, _FLOAT_COUNTERS[name]] = int(np.float32(value).view(np.int32))
        else:
            object.__setattr__(self, name, value)

    # ---- observations (CyberDefenseEnv.py:146/241/194) ----
    def _obs(self, mode, dtype):
        o = self._venv.observe(mode)
        torch.cuda.synchronize()
        return o[0].cpu().numpy().astype(dtype)

    def _get_state(self):
        return self._obs(3, np.float64)

    def _get_defender_state(self):
        return self._obs(1, np.float64)

    def _get_attacker_state(self):
        return self._obs(2, np.float32)

    # ---- actions ----
    def _mode_id(self):
        if self.mode not in ("defender", "attacker"):
            raise ValueError("Invalid mode")
        return 1 if self.mode == "attacker" else 0

    def sample_action(self):
        ab = self._venv.sample_actions(self._mode_id())
        torch.cuda.synchronize()
        hdr = ab.hdr[0].cpu().numpy().view(np.uint32)
        mask = ab.mask[0].cpu().numpy().view(np.uint32)
        devs = [d for d in range(self._net.M) if (mask[d >> 5] >> (d & 31)) & 1]
        return (int(np.int8(hdr[0] & 0xFF)), np.array([int(np.int8(hdr[1] & 0xFF))], dtype=int), devs, int(np.int32(hdr[3])))

    def randomize_compromise_and_ownership(self):
        self._venv.randomize_compromise_and_ownership()

    def _rebuild_graph_cache(self):
        """No-op: the kernels read the network tables directly (volt_typhoon_env.py:456-483)."""

    def _pack(self, action, mode):
        if action is not None:
            atype, ex, devs, app = action
            devs = [int(d) for d in devs]
            if int(atype) in (11, 12, 13) and mode == 0 and len(devs) == 0 and self._base_line == "Nash":
                raise ValueError("device_indices[0] is required")  # volt:965-966, :1103-1104, :1112-1113
            order_form = devs != sorted(set(devs))
            action = (atype, ex, devs, app)
        else:
            order_form = False
        return ActionBatch.pack([action], mode, self._net.M, order_form=order_form)

    def _finish(self, raw, shaped, done, action, executed):
        torch.cuda.synchronize()
        pre = self._venv.pre_masks()[0].cpu().numpy().view(np.uint32)
        M = self._net.M
        bits = lambda row: np.array([(pre[row, d >> 5] >> (d & 31)) & 1 for d in range(M)], np.float64)
        st = np.zeros((M, 6), np.float64)  # the pre-evolve `state` step() returns (volt:1306)
        st[:, 0], st[:, 1] = self._net.os_val, self._net.ver_val
        st[:, 2], st[:, 4], st[:, 5] = bits(0), bits(1), bits(2)
        self.state = st.reshape(-1)
        info = {
            "mode": self.mode, "step_count": self.step_num - 1 if executed else self.step_num,
            "revert_count": self.revert_count, "checkpoint_count": self.checkpoint_count,
            "defensive_cost": self.defensive_cost, "clearning_cost": self.clearning_cost, "Scan_count": self.scan_cnt,
            "action_taken": action, "work_done": self.work_done, "Compromised_devices": self.compromised_devices_cnt,
            "Edges Blocked": self.edges_blocked, "Edges Added": self.edges_added,
        }
        logs = []  # the hop log itself is not materialised; its length drives the kernel (CYG_S_LOGS)
        return self.state, float(raw[0].item()), float(shaped[0].item()), bool(done[0].item()), info, logs

    def step(self, action, agent_cnt=None):
        if isinstance(action, (list, tuple)) and action and isinstance(action[0], (list, tuple)):
            return self.step_grouped(action)
        mode = self._mode_id()
        hdr, mask, order = self._pack(action, mode)
        flags = 0
        if agent_cnt is not None and agent_cnt != self._net.M:
            flags |= K.STEP_SKIP_WORK  # volt:1207, :1307
        raw, shaped, done = self._venv.step(self._venv.to_device(hdr, mask, order), flags=flags, want_pre=True)
        return self._finish(raw, shaped, done, action, executed=not flags)

    def step_grouped(self, groups):
        mode = self._mode_id()
        batches = []
        any_order = any(list(map(int, g[2])) != sorted(set(map(int, g[2]))) for g in groups)
        for g in groups:
            h, m, o = ActionBatch.pack([(g[0], g[1], [int(d) for d in g[2]], g[3])], mode, self._net.M, order_form=any_order)
            batches.append(self._venv.to_device(h, m, o))
        raw, shaped, done = self._venv.step_grouped(batches, want_pre=True)
        return self._finish(raw, shaped, done, list(groups), executed=True)


CyberDefenseEnv = Volt_Typhoon_CyberDefenseEnv
