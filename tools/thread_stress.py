"""Repeats tests/test_gpu_threads_devlists.py::test_five_threads_trace_on_one_visualization in one process and prints every
failure (which lane-call differed and where), for chasing rare interleavings.  python tools/thread_stress.py [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    from galaxy_b200 import gpu
    from tests import test_gpu_threads_devlists as t
    bad = 0
    for i in range(reps):
        try:
            t.test_five_threads_trace_on_one_visualization.__wrapped__(gpu) if hasattr(
                t.test_five_threads_trace_on_one_visualization, "__wrapped__") else t.test_five_threads_trace_on_one_visualization(gpu)
        except AssertionError as e:
            bad += 1
            print("rep %d FAILED: %s" % (i, str(e)[:1500]), flush=True)
    print("thread_stress: %d of %d repetitions failed" % (bad, reps))


if __name__ == "__main__":
    main()
