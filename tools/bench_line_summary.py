#!/usr/bin/env python
"""One line per bench.py run for tools/first_gpu_call.sh: ms/step, it/s and the per-stage split of the last JSON line of a log."""
import json
import sys


def summarise(path, name):
    try:
        line = [l for l in open(path) if l.startswith("{")][-1]
        j = json.loads(line)
        st = j["roofline"]["stage_ms"]
        clk = j.get("clocks") or {}
        return (f"    {name:28s} {j['ms_per_step']:.4f} ms/step  {j['iters_per_s']:8.1f} it/s | " + "  ".join(f"{k}={v:.4f}" for k, v in st.items())
                + f" | e2e {j.get('e2e', {}).get('value', 0) / max(j.get('value', 1), 1e-9):.3f} of value | sm {clk.get('sm_mhz')} MHz {clk.get('reasons')}")
    except Exception as e:                                    # noqa: BLE001 -- a summary line must never fail the script
        return f"    {name}: no JSON line ({e})"


if __name__ == "__main__":
    print(summarise(sys.argv[1], sys.argv[2]))
