"""Developer tool: run a timing probe once per library variant (tools/build_variants.py), each in its own process.

    python tools/time_variants.py tools/split_time.py base d8 ...
"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
lib = os.path.join(os.path.dirname(here), "pytestbeam_b200", "_lib")
probe = sys.argv[1]
for name in sys.argv[2:]:
    env = dict(os.environ, PTB_LIB_PATH=os.path.join(lib, f"var_{name}.so"))
    print(f"==== {name}", flush=True)
    subprocess.run([sys.executable, probe], env=env)
