# timing-knob sweep: bash tools/knobs.sh VAR v1 v2 ...   (prints the per-class step table for each value of env VAR)
var=$1; shift
for d in "$@"; do
  echo "== $var $d"
  env $var=$d python tools/step_table.py 2>&1 | grep -E "forward|conv3x3"
done
