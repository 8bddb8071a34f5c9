"""Prints the constructor stages of a bench line (stdin or file): setup_s | value | e2e | stage increments."""
import json, sys
d = json.loads((open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin).read())
prev, out = 0.0, []
for name, t in d["cold"]["setup_log"]:
    out.append("%s +%.2f" % (name.split(" ")[0], t - prev))
    prev = t
print(round(d["setup_s"], 2), round(d["value"], 1), round(d["e2e"]["value"], 1), "|", ", ".join(out))
