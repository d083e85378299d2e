"""Wall-clock guard against stuck device work (a collective whose peer never arrives, a kernel
spinning on a flag): the reference has no multi-device path, so it has no counterpart there.

    guard = HangGuard(120, who="rank 3")
    guard.beat("warm-up step 2")      # before every phase that may block
    ...
    guard.stop()

If no beat arrives for `timeout_s` seconds the guard prints which phase stalled, dumps the Python
stack of every thread (faulthandler) to stderr and ends the process with exit code 3 -- under
torchrun that tears the other ranks down as well, so a hang costs `timeout_s`, not the NCCL
watchdog's ten minutes, and leaves a diagnosis behind.
"""
import faulthandler
import os
import sys
import threading
import time


class HangGuard:
    def __init__(self, timeout_s, who=""):
        self.timeout_s = float(timeout_s)
        self.who = who
        self.phase = "start"
        self.t_last = time.monotonic()
        self._stop = threading.Event()
        self.thread = None
        if self.timeout_s > 0:
            self.thread = threading.Thread(target=self._run, name="hang-guard", daemon=True)
            self.thread.start()

    def beat(self, phase=None):
        if phase is not None:
            self.phase = phase
        self.t_last = time.monotonic()

    def stop(self):
        self._stop.set()

    def _run(self):
        while not self._stop.wait(1.0):
            idle = time.monotonic() - self.t_last
            if idle > self.timeout_s:
                sys.stderr.write(f"\n[hang-guard] {self.who}: no progress for {idle:.0f} s in phase "
                                 f"'{self.phase}'; Python stacks of all threads follow, then exit(3)\n")
                sys.stderr.flush()
                try:
                    faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                    sys.stderr.flush()
                finally:
                    os._exit(3)
