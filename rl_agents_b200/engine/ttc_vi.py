"""ValueIterationAgent on HighwayLite scenes, batched on the device (b2_highway_ttc_vi): per scene the time-to-collision
grid MDP of `env.unwrapped.to_finite_mdp()` is built and the agent's fixed point (value_iteration.py:42-73) is iterated by
one warp in shared memory.  Torch tensors own the buffers; nothing is computed on the host."""
import numpy as np

from rl_agents_b200 import _lib

TTC_STATES, TTC_ACTIONS = 120, 5


class HighwayTTCVI(object):
    def __init__(self, gamma=1.0, iterations=100, rtol=1e-5, atol=1e-8, device="cuda"):
        import torch
        self.lib = _lib.load()
        self.gamma, self.iterations, self.rtol, self.atol = float(gamma), int(iterations), float(rtol), float(atol)
        self.device = torch.device(device)

    def solve(self, scenes, want_q=True):
        """scenes: int32 [n, 136] (torch CUDA tensor or numpy).  Returns dict(action [n], state [n], sweeps [n],
        q [n, 120, 5] or None) of torch tensors on the device."""
        import torch
        if not torch.is_tensor(scenes):
            scenes = torch.from_numpy(np.ascontiguousarray(scenes, dtype=np.int32))
        scenes = scenes.to(self.device, dtype=torch.int32).contiguous().reshape(-1, _lib.HW_STATE_WORDS)
        n = scenes.shape[0]
        q = torch.empty((n, TTC_STATES, TTC_ACTIONS), dtype=torch.float64, device=self.device) if want_q else None
        action = torch.empty(n, dtype=torch.int32, device=self.device)
        state = torch.empty(n, dtype=torch.int32, device=self.device)
        sweeps = torch.empty(n, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.b2_highway_ttc_vi(_lib.ptr(scenes), n, self.gamma, self.iterations, self.rtol, self.atol,
                                              _lib.ptr(q) if want_q else None, _lib.ptr(action), _lib.ptr(state),
                                              _lib.ptr(sweeps), _lib.current_stream()))
        return {"action": action, "state": state, "sweeps": sweeps, "q": q}
