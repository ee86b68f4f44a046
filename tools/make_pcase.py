import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hopefoam_b200 import meshgen
from tests.case_writer import write_euler_case
n = int(sys.argv[1]) if len(sys.argv) > 1 else 354
mg = meshgen.jittered_square(n)
dt = 0.09 / n
write_euler_case("/tmp/pcase", mg, 4, dt, dt * 30)
print("pcase written")
