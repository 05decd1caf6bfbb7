"""
Host hooks of the time loop: commands and state save / load.

These are not compute components; the device runner only has to honour them
("Commands", "State save/load").  Names, attributes and
firing rules follow the reference so user scripts run unchanged:
  Command / CommandSequence   finitewave/core/command/{command,command_sequence}.py
  StateSaver / StateSaverCollection / StateLoader
                              finitewave/core/state/{state_saver,state_loader}.py
The model synchronises ``model.__dict__[var]`` with the device around every hook
(model.CardiacModel.run), so ``execute(model)`` / ``save()`` see plain numpy arrays.

Checkpoints taken DURING a run (a StateSaver whose ``time`` lies before ``t_max``) do not
stall the step loop: the model snapshots the device state on the compute
stream, a copy stream moves the snapshot into pinned host memory, and a writer thread
stores the ``.npy`` files (``AsyncCheckpointWriter``); the files are complete when ``run()``
returns.  The final checkpoint (``time = -1``) is written from the arrays ``run()`` downloads
anyway.
"""
import queue
import threading
from pathlib import Path

import numpy as np


class Command:
    """User callback fired once, after the step on which ``model.t >= time``."""

    def __init__(self, time=None):
        self.t = time
        self.passed = False

    def execute(self, model):
        raise NotImplementedError

    def update_status(self, model):
        self.passed = model.t >= self.t
        return self.passed


class CommandSequence:
    def __init__(self):
        self.sequence = []
        self.model = None

    def initialize(self, model):
        self.model = model
        for command in self.sequence:
            command.passed = False

    def add_command(self, command):
        self.sequence.append(command)

    def remove_commands(self):
        self.sequence = []

    def execute_next(self):
        for command in self.sequence:
            if not command.passed and command.update_status(self.model):
                command.execute(self.model)


class StateSaver:
    """Writes every ``model.state_vars`` array to ``<path>/<var>.npy`` once, at
    ``time`` (or at ``t_max`` when ``time < 0``)."""

    def __init__(self, path=".", time=-1):
        self.path = path
        self.passed = False
        self.model = None
        self.time = time

    def initialize(self, model):
        self.model = model
        self.passed = self.path == ""

    def due(self):
        if self.passed:
            return False
        if self.time < 0:
            return self.model.t >= self.model.t_max
        return self.model.t >= self.time

    def save(self):
        if not self.due():
            return
        Path(self.path).mkdir(parents=True, exist_ok=True)
        for var in self.model.state_vars:
            self._save_variable(Path(self.path).joinpath(var + ".npy"),
                                self.model.__dict__[var])
        self.passed = True

    def _save_variable(self, var_path, var):
        np.save(var_path, var)


class AsyncCheckpointWriter:
    """Writer thread(s) of the checkpoints taken while the simulation keeps running."""

    def __init__(self):
        self.threads = []
        self.errors = []

    def submit(self, path, names, arrays, done_event, keepalive):
        def work():
            try:
                done_event.synchronize()            # the device-to-host copies have landed
                Path(path).mkdir(parents=True, exist_ok=True)
                for name, a in zip(names, arrays):
                    np.save(Path(path).joinpath(name + ".npy"), a)
            except Exception as e:                  # re-raised by wait()
                self.errors.append(e)
            finally:
                keepalive.clear()
        th = threading.Thread(target=work, daemon=True)
        th.start()
        self.threads.append(th)

    def wait(self):
        for th in self.threads:
            th.join()
        self.threads = []
        if self.errors:
            e, self.errors = self.errors[0], []
            raise e


class FrameStreamer:
    """Streams device arrays to ``.npy`` files while the simulation keeps running (frame
    dumps of the animation trackers).  Per frame: a device snapshot on
    the compute stream (taken by the caller), a device-to-host copy on a copy stream into
    one of ``depth`` pinned buffers, and the file write on a writer thread.  ``submit``
    only blocks when all pinned buffers are still being written (back-pressure)."""

    def __init__(self, shape, depth=3):
        import torch
        self.torch = torch
        self.shape = tuple(shape)
        self.copy_stream = torch.cuda.Stream()
        self.free = queue.Queue()
        for _ in range(depth):
            self.free.put(torch.empty(self.shape, dtype=torch.float64, pin_memory=True))
        self.jobs = queue.Queue()
        self.errors = []
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        while True:
            job = self.jobs.get()
            if job is None:
                return
            host, done, path, dtype, select = job
            try:
                done.synchronize()
                a = host.numpy()
                if select is not None:
                    a = a[select]
                np.save(path, a.astype(dtype))
            except Exception as e:
                self.errors.append(e)
            finally:
                self.free.put(host)
                self.jobs.task_done()

    def submit(self, snapshot, path, dtype, select=None):
        """snapshot: a device tensor nobody will overwrite (its producer ran on the current
        stream); select: optional index applied on the host (a slice of a 3D frame)."""
        torch = self.torch
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        host = self.free.get()                       # back-pressure
        done = torch.cuda.Event()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            host.copy_(snapshot, non_blocking=True)
            snapshot.record_stream(self.copy_stream)
            done.record(self.copy_stream)
        self.jobs.put((host, done, path, dtype, select))

    def wait(self):
        self.jobs.join()
        if self.errors:
            e, self.errors = self.errors[0], []
            raise e

    def close(self):
        self.wait()
        self.jobs.put(None)


class StateSaverCollection(StateSaver):
    def __init__(self):
        super().__init__()
        self.savers = []

    def initialize(self, model):
        self.model = model
        for saver in self.savers:
            saver.initialize(model)

    def due(self):
        return any(s.due() for s in self.savers)

    def save(self):
        for saver in self.savers:
            saver.save()


class StateLoader:
    """Reads ``<path>/<var>.npy`` into ``model.<var>`` once at the start of run()."""

    def __init__(self, path=""):
        self.path = path
        self.passed = True
        self.model = None

    def initialize(self, model):
        self.model = model
        self.passed = self.path == ""
        if not Path(self.path).exists():
            raise FileNotFoundError(f"Unable to load state from {self.path}. "
                                    "Directory does not exist.")

    def load(self):
        if self.passed:
            return
        for var in self.model.state_vars:
            setattr(self.model, var, self._load_variable(Path(self.path).joinpath(var + ".npy")))
        self.passed = True

    def _load_variable(self, var_path):
        return np.load(var_path)
