"""Per-kernel durations of the LAST repetition in an ncu launch list (ncu --metrics gpu__time_duration.sum --csv):
    python tools/launch_breakdown.py gpurun_out/launches.csv [repetitions]"""
import csv
import sys

path, reps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(x["Kernel Name"][:64], float(x["Metric Value"].replace(",", "")) / 1000) for x in csv.DictReader(lines)]
n = len(rows) // reps
tot = 0.0
for k, v in rows[-n:]:
    print(f"{v:9.1f} us  {k}")
    tot += v
print(f"{tot:9.1f} us  total ({n} launches)")
