"""Run the reference's own CUDA DNN build (oracle/_ref/jref_cuda) on one short utterance of a DNN workload and show what
it says -- the bench leg runs it with the log muted."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from julius_b200 import workload

name = sys.argv[1] if len(sys.argv) > 1 else "dnn20k"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
m = workload.synth_model(name)
fn = "/tmp/refcuda_u0.mfc"
workload.write_input(name, fn, workload.sample_inputs(name, m, 1, T, seed=5)[0])
args = [os.path.join(ROOT, "oracle", "_ref", "jref_cuda"), "-dump", "/dev/null"] + workload.ref_args(name)
p = subprocess.run(args, input=fn + "\n" + fn + "\n", text=True, capture_output=True, env=dict(os.environ, JREF_PER_UTT="1"), timeout=240)
print("rc", p.returncode)
print("\n".join(p.stdout.splitlines()[-25:]))
print("STDERR", p.stderr[-1500:])
