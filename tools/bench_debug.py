import faulthandler, sys, runpy, os
faulthandler.dump_traceback_later(240, exit=True)
sys.argv = ["bench.py", "--gpus", os.environ.get("WORLD_SIZE", "1"), "--steps", "5", "--warmup", "3", "--points-per-gpu", os.environ.get("PPG", "1000000")]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bench.py"), run_name="__main__")
