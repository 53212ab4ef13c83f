"""Host-buffer stepping: the call a CPU-side RL loop makes (actions in host memory, observations /
rewards / dones wanted back in host memory), as the reference's gym adapters do after `jax.device_get`
(/root/reference/po_brax/envs/wrappers.py:126-172 convert to / from host arrays every step).

The batch is split into independent sub-batches (envs never interact, so this is exact -- see the
shard-equivalence test), each with its own stream, so the host->device copy of chunk c+1 overlaps the
step kernel of chunk c and the device->host copy of chunk c-1."""
import os
from typing import List, Optional

import torch

from .envs.env import Env


def _gpu_numa_node(index: int) -> Optional[int]:
    """NUMA node of CUDA device `index` from sysfs (None when the platform does not say: containers often report -1)."""
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(index), 'pci_domain_id', 0)
        dev = getattr(torch.cuda.get_device_properties(index), 'pci_device_id', 0)
        path = f'/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node'
        node = int(open(path).read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(index: int) -> dict:
    """One process per GPU: restrict this process to the CPUs of its GPU's NUMA node BEFORE the pinned host buffers
    are allocated, so first touch puts them next to the PCIe root the GPU hangs off (8 ranks pinning memory wherever
    the allocator lands send half of the device->host traffic across the socket interconnect). Best effort: returns
    what was done ({'node': ..., 'cpus': n} or {'node': None, 'why': ...}); never raises."""
    node = _gpu_numa_node(index)
    if node is None:
        return {'node': None, 'why': 'sysfs reports no NUMA node for the GPU'}
    try:
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {'node': node, 'why': 'no allowed CPU on that node'}
        os.sched_setaffinity(0, cpus)
        return {'node': node, 'cpus': len(cpus)}
    except Exception as e:   # noqa: BLE001
        return {'node': node, 'why': f'{type(e).__name__}: {e}'}


class HostStepper:
    def __init__(self, env_name: str, batch_size: int, chunks: int = 4, device=None, **create_kwargs):
        if batch_size % chunks:
            raise ValueError('batch_size must be divisible by chunks')
        self.n, self.chunks, self.m = batch_size, chunks, batch_size // chunks
        self.envs: List[Env] = [Env(env_name, batch_size=self.m, device=device, **create_kwargs) for _ in range(chunks)]
        self.device = self.envs[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(chunks)]
        d, a = self.envs[0].observation_size, self.envs[0].action_size
        pin = dict(pin_memory=True)
        self.action_host = torch.zeros((batch_size, a), dtype=torch.float32, **pin)
        self.obs_host = torch.zeros((batch_size, d), dtype=torch.float32, **pin)
        self.reward_host = torch.zeros(batch_size, dtype=torch.float32, **pin)
        self.done_host = torch.zeros(batch_size, dtype=torch.float32, **pin)
        self._act_dev = [torch.empty((self.m, a), dtype=torch.float32, device=self.device) for _ in range(chunks)]
        self.states = [None] * chunks
        self.h2d_bytes_per_step = self.action_host.numel() * 4
        self.d2h_bytes_per_step = (self.obs_host.numel() + self.reward_host.numel() + self.done_host.numel()) * 4

    def _order_after_current(self):
        """Inputs handed to reset / step may have been produced on the caller's current stream (device keys from
        shard_keys, an action tensor): every chunk stream first waits for the work enqueued there so far."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)

    def reset(self, keys) -> torch.Tensor:
        """keys: uint32 [N, 2] (host array or device tensor). Returns the host observation buffer."""
        self._order_after_current()
        for c, (env, st) in enumerate(zip(self.envs, self.streams)):
            sl = slice(c * self.m, (c + 1) * self.m)
            with torch.cuda.stream(st):
                k = keys[sl]
                if isinstance(k, torch.Tensor) and k.is_cuda:
                    k.record_stream(st)   # allocated on the caller's stream, consumed on this one
                s = env.reset(k)
                self.states[c] = s
                self.obs_host[sl].copy_(s.obs, non_blocking=True)
        self.sync()
        return self.obs_host

    def step_async(self):
        """Enqueue: action_host -> device, fused step, obs/reward/done -> host, per chunk on its stream."""
        self._order_after_current()
        for c, (env, st) in enumerate(zip(self.envs, self.streams)):
            sl = slice(c * self.m, (c + 1) * self.m)
            with torch.cuda.stream(st):
                self._act_dev[c].copy_(self.action_host[sl], non_blocking=True)
                s = env.step(self.states[c], self._act_dev[c])
                self.states[c] = s
                self.obs_host[sl].copy_(s.obs, non_blocking=True)
                self.reward_host[sl].copy_(s.reward, non_blocking=True)
                self.done_host[sl].copy_(s.done, non_blocking=True)

    def sync(self):
        for st in self.streams:
            st.synchronize()

    def step(self, action_host=None):
        """gym-style: returns (obs, reward, done) host tensors after the step has completed."""
        if action_host is not None:
            self.action_host.copy_(action_host)
        self.step_async()
        self.sync()
        return self.obs_host, self.reward_host, self.done_host
