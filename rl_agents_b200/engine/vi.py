"""Value-iteration engine (device side of ValueIterationAgent)."""
import numpy as np

from rl_agents_b200 import _lib

MODES = {"deterministic": _lib.VI_DETERMINISTIC, "stochastic": _lib.VI_STOCHASTIC, "sparse": _lib.VI_SPARSE}


class VIEngine(object):
    """Holds the slab [row_begin, row_end) of an MDP's tables on one device and
    runs Bellman sweeps (b2_vi_sweep).  With world_size > 1 every rank owns a
    slab and V is all-gathered after each sweep (rl_agents_b200.distributed)."""

    def __init__(self, mode, transition, reward, terminal, nxt=None, gamma=1.0, device="cuda",
                 row_begin=0, row_end=None, n_states=None, rtol=1e-5, atol=1e-8):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.mode = mode

        def dev(x, dtype):
            """host arrays are uploaded; tensors already on the device (slabs generated / loaded in place,
            rl_agents_b200.distributed tables_are_local) are used as they are."""
            if isinstance(x, torch.Tensor):
                return x.to(device=self.device, dtype=dtype).contiguous()
            np_dtype = {torch.int32: np.int32, torch.float64: np.float64, torch.uint8: np.uint8}[dtype]
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np_dtype), device=self.device)
        reward = dev(reward, torch.float64)
        rows, A = reward.shape
        self.n_actions = A
        self.n_states = int(n_states if n_states is not None else rows)
        self.row_begin = int(row_begin)
        self.row_end = int(row_end if row_end is not None else self.row_begin + rows)
        assert self.row_end - self.row_begin == rows
        if mode == "deterministic":
            self.transition = dev(transition, torch.int32)
            self.next, self.n_next = None, 1
        elif mode == "stochastic":
            self.transition = dev(transition, torch.float64)
            self.next, self.n_next = None, self.n_states
        elif mode == "sparse":
            self.transition = dev(transition, torch.float64)
            self.next = dev(nxt, torch.int32)
            self.n_next = int(self.next.shape[-1])
        else:
            raise ValueError("Unknown mode")
        self.reward = reward
        self.terminal = dev(terminal, torch.uint8)
        self.problem = _lib.VIProblem(MODES[mode], A, self.n_next, 0, self.n_states, self.row_begin, self.row_end,
                                      float(gamma), rtol, atol, self.transition.data_ptr(),
                                      self.next.data_ptr() if self.next is not None else None,
                                      self.reward.data_ptr(), self.terminal.data_ptr())
        self.q = [torch.zeros(rows, A, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.v = [torch.zeros(self.n_states, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.viol = None

    def bytes_per_sweep(self):
        """Algorithmic HBM bytes of one sweep (SURVEY 8d formula, + the Q_old read)."""
        rows, A, B = self.row_end - self.row_begin, self.n_actions, self.n_next
        if self.mode == "deterministic":
            return rows * A * (4 + 8 + 8 + 8 + 8) + rows * 9
        return rows * A * B * (8 + (4 if self.mode == "sparse" else 0) + 8) + rows * A * (8 + 8 + 8) + rows * 9

    def reset(self, iterations):
        for t in self.q + self.v:
            t.zero_()
        self.viol = self.torch.zeros(max(int(iterations), 1), dtype=self.torch.int32, device=self.device)

    def sweep(self, k):
        """Enqueue sweep k: reads q[k%2], v[k%2]; writes q[(k+1)%2], v[(k+1)%2]."""
        _lib.check(self.lib.b2_vi_sweep(self.problem, _lib.ptr(self.v[k & 1]), _lib.ptr(self.q[k & 1]),
                                        _lib.ptr(self.q[(k + 1) & 1]), _lib.ptr(self.v[(k + 1) & 1]),
                                        _lib.ptr(self.viol), k, _lib.current_stream()))

    def solve(self, iterations):
        """fixed_point_iteration (value_iteration.py:65-73) without host round trips;
        returns (Q tensor on device, sweeps performed)."""
        self.reset(iterations)
        _lib.check(self.lib.b2_vi_solve(self.problem, _lib.ptr(self.q[0]), _lib.ptr(self.q[1]), _lib.ptr(self.v[0]),
                                        _lib.ptr(self.v[1]), _lib.ptr(self.viol), int(iterations),
                                        _lib.current_stream()))
        return self.result(iterations)

    def result(self, iterations):
        viol = self.viol[:iterations].cpu().numpy()     # synchronises
        zero = np.nonzero(viol == 0)[0]
        if zero.size:                                    # converged at sweep k: return the OLD iterate
            k = int(zero[0])
            return self.q[k & 1], k + 1
        return self.q[iterations & 1], int(iterations)


class RobustVIEngine(object):
    """Device side of RobustValueIterationAgent: M models, Q <- min_m (R_m + gamma E_m[max_a Q])."""

    def __init__(self, mode, transitions, rewards, gamma=1.0, device="cuda", rtol=1e-5, atol=1e-8):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        rewards = np.asarray(rewards, dtype=np.float64)
        self.n_models, self.n_states, self.n_actions = rewards.shape
        if mode == "deterministic":
            self.transition = torch.as_tensor(np.ascontiguousarray(transitions, dtype=np.int32), device=self.device)
        elif mode == "stochastic":
            self.transition = torch.as_tensor(np.ascontiguousarray(transitions, dtype=np.float64), device=self.device)
        else:
            raise ValueError("Unknown mode")
        self.reward = torch.as_tensor(np.ascontiguousarray(rewards), device=self.device)
        self.problem = _lib.VIProblem(MODES[mode], self.n_actions, self.n_states, 0, self.n_states, 0, self.n_states,
                                      float(gamma), rtol, atol, self.transition.data_ptr(), None,
                                      self.reward.data_ptr(), None)
        self.q = [torch.zeros(self.n_states, self.n_actions, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.v = [torch.zeros(self.n_states, dtype=torch.float64, device=self.device) for _ in range(2)]

    def solve(self, iterations):
        for t in self.q + self.v:
            t.zero_()
        self.viol = self.torch.zeros(max(int(iterations), 1), dtype=self.torch.int32, device=self.device)
        for k in range(int(iterations)):
            _lib.check(self.lib.b2_vi_robust_sweep(self.problem, self.n_models, _lib.ptr(self.v[k & 1]),
                                                   _lib.ptr(self.q[k & 1]), _lib.ptr(self.q[(k + 1) & 1]),
                                                   _lib.ptr(self.v[(k + 1) & 1]), _lib.ptr(self.viol), k,
                                                   _lib.current_stream()))
        viol = self.viol[:iterations].cpu().numpy()
        zero = np.nonzero(viol == 0)[0]
        if zero.size:
            k = int(zero[0])
            return self.q[k & 1], k + 1
        return self.q[iterations & 1], int(iterations)
