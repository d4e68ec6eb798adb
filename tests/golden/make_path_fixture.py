"""Regenerates tests/golden/path_yaml.json from the reference's only golden artefact,
graph_core/tests/path.yaml (14 waypoints of one historical RRT* run in the Cube3d world).
Run in the build container, where /root/reference exists; the JSON is what travels."""
import json
import re
import sys
from pathlib import Path

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/graph_core/tests/path.yaml")
rows = [[float(v) for v in m.group(1).split(",")] for m in re.finditer(r"-\s*\[([^\]]+)\]", src.read_text())]
out = Path(__file__).parent / "path_yaml.json"
out.write_text(json.dumps({"source": "graph_core/tests/path.yaml", "waypoints": rows}, indent=1))
print(len(rows), "waypoints ->", out)
