"""Solver steps on HOST-resident states, with the PCIe copies overlapped with the compute.

The reference-facing call `lib.simulate(mconf, batch_dict, net, method)` (pytorch/lib/simulate.py:28-171) works on
device tensors.  A caller whose states live in host memory (one request per step: a service answering "advance this
state", a post-processing chain that wants every step on the host) pays, per step, the upload of the state and the
download of the result: at 4096^2 that is 268 MB in + 268 MB out = ~10 ms of PCIe against 1.9 ms of kernels.  PCIe is
full duplex and the copy engines run beside the SMs, so the three legs of CONSECUTIVE INDEPENDENT steps can run at
once:

        copy-in stream   : H2D(k+1)
        compute stream   :            step(k)
        copy-out stream  :                        D2H(k-1)

`HostStepPipeline.submit(host_in, host_out)` enqueues one step and returns; `flush()` waits for everything
submitted.  Every step still uploads its own inputs and downloads its own results -- only the waiting is shared.
What is NOT uploaded: `p`.  Neither projection reads the incoming pressure (Jacobi starts from p = 0,
fluids_init.cpp:858-1003; the shipped ScaleNet configuration has inputChannels.pDiv off, model.py:109-111), the
step only writes it; the slot's `p` buffer is allocated once.

Ordering (events, no host synchronisation inside submit):
  * slot k's input buffers are rewritten only after the step that read them has finished (compute_done[k]);
  * the compute stream waits for the slot's upload (h2d_done[k]) and -- before the tensors of the slot's previous
    results are released to the allocator -- for their download (d2h_done[k]);
  * the copy-out stream waits for the step (compute_done[k]).
"""
import importlib

import torch


class _Slot:
    def __init__(self, like, device):
        self.inp = {k: torch.empty(like[k].shape, dtype=torch.float32, device=device)
                    for k in ('U', 'flags', 'density')}
        self.inp['p'] = torch.zeros(like['p'].shape, dtype=torch.float32, device=device)
        self.out = None
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.used = False


class HostStepPipeline:
    """`depth` slots of device buffers cycling through upload -> step -> download.

        pipe = HostStepPipeline(mconf, net, 'jacobi', like=host_state, device='cuda:0', masks=bc_tensors)
        for state_in, state_out in work:          # pinned host tensors: p, U, flags, density / p, U, density
            pipe.submit(state_in, state_out)
        pipe.flush()                              # state_out of every submitted step is complete

    `masks`: the device-resident imposed-value tensors of the simulation (UBC, UBCInvMask, densityBC,
    densityBCInvMask), shared by all steps.  Host tensors must be pinned for the copies to be asynchronous (pageable
    memory works but serialises)."""

    def __init__(self, mconf, net, sim_method, like, device, masks=None, depth=2):
        assert depth >= 2, "a pipeline needs at least two slots"
        self.sim = importlib.import_module(__package__ + ".simulate")
        self.mconf, self.net, self.method = mconf, net, sim_method
        self.device = torch.device(device)
        self.masks = dict(masks or {})
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.slots = [_Slot(like, self.device) for _ in range(depth)]
        self.n = 0
        self.h2d_bytes_per_step = sum(like[k].numel() * 4 for k in ('U', 'flags', 'density'))
        self.d2h_bytes_per_step = sum(like[k].numel() * 4 for k in ('p', 'U', 'density'))

    def submit(self, host_in, host_out):
        slot = self.slots[self.n % len(self.slots)]
        cur = torch.cuda.current_stream(self.device)
        if slot.used:
            self.s_in.wait_event(slot.compute_done)        # the step that read these buffers is done
        with torch.cuda.stream(self.s_in):
            for k in ('U', 'flags', 'density'):
                slot.inp[k].copy_(host_in[k], non_blocking=True)
            slot.h2d_done.record(self.s_in)
        cur.wait_event(slot.h2d_done)
        if slot.used:
            cur.wait_event(slot.d2h_done)                  # previous results of this slot have left the device
        slot.out = None                                    # ... only now may the allocator reuse them
        d = dict(slot.inp)
        d.update(self.masks)
        with torch.no_grad():
            self.sim.simulate(self.mconf, d, self.net, self.method)
        slot.out = {k: d[k] for k in ('p', 'U', 'density')}
        slot.compute_done.record(cur)
        self.s_out.wait_event(slot.compute_done)
        with torch.cuda.stream(self.s_out):
            for k in ('p', 'U', 'density'):
                host_out[k].copy_(slot.out[k], non_blocking=True)
            slot.d2h_done.record(self.s_out)
        slot.used = True
        self.n += 1

    def flush(self):
        self.s_out.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
