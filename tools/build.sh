#!/bin/bash
# always builds from the repo root, prints errors only, then the library timestamp
cd "$(dirname "$0")/.." && python -c "import __graft_entry__ as g; g.build(force=True)" 2>&1 | grep -iE "error|warning: v" -A4 | head -40
ls -la --time-style=+%T spark-tfrecord_b200/libtfrgpu.so | awk '{print "built", $6}'
