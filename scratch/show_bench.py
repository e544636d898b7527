import json, sys
d = json.load(open(sys.argv[1]))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "launch_ms", d["roofline"]["avg_launch_ms"])
print("e2e", d["e2e"])
print("cpu", d["cpu_baseline"])
print("paths", d["path_taken"], "launches", d["gpu_launches"], "host_ms", d["host_enqueue_ms_per_step"], "clocks", d["clocks"])
print(json.dumps(d["other_configs"], indent=1)[:7000])
