import os
import sys
import time
from collections import abc


def is_str(x):
    return isinstance(x, str)


def is_seq_of(seq, expected_type, seq_type=None):
    exp_seq_type = abc.Sequence if seq_type is None else seq_type
    if not isinstance(seq, exp_seq_type):
        return False
    return all(isinstance(item, expected_type) for item in seq)


def is_list_of(seq, expected_type):
    return is_seq_of(seq, expected_type, seq_type=list)


def is_tuple_of(seq, expected_type):
    return is_seq_of(seq, expected_type, seq_type=tuple)


def mkdir_or_exist(dir_name, mode=0o777):
    if dir_name == "":
        return
    os.makedirs(os.path.expanduser(dir_name), mode=mode, exist_ok=True)


class ProgressBar:
    """Text progress bar with the `update()` call mogen/apis/test.py:18,62 uses."""

    def __init__(self, task_num=0, bar_width=50, start=True, file=sys.stdout):
        self.task_num, self.bar_width, self.completed, self.file = task_num, bar_width, 0, file
        if start:
            self.start()

    def start(self):
        self.start_time = time.time()
        self.file.write(f"[{' ' * self.bar_width}] 0/{self.task_num}, elapsed: 0s\n" if self.task_num > 0 else "completed: 0\n")
        self.file.flush()

    def update(self, num_tasks=1):
        self.completed += num_tasks
        elapsed = time.time() - self.start_time
        if self.task_num > 0:
            frac = min(1.0, self.completed / float(self.task_num))
            done = int(self.bar_width * frac)
            self.file.write(f"\r[{'>' * done}{' ' * (self.bar_width - done)}] {self.completed}/{self.task_num}, "
                            f"{self.completed / max(elapsed, 1e-9):.1f} task/s, elapsed: {int(elapsed + 0.5)}s")
        else:
            self.file.write(f"\rcompleted: {self.completed}, elapsed: {int(elapsed + 0.5)}s")
        self.file.flush()
