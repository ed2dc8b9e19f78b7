"""Wall-clock timer with named accumulators (reference: distance3d/benchmark.py:5-22)."""
import time


class Timer:
    def __init__(self):
        self.start_times_ = {}
        self.total_time_ = {}

    def start(self, name):
        self.start_times_[name] = time.perf_counter()

    def stop(self, name):
        return time.perf_counter() - self.start_times_.pop(name)

    def stop_and_add_to_total(self, name):
        self.total_time_[name] = self.total_time_.get(name, 0.0) + self.stop(name)
