# same-box A/B of several builds: bash tools/ab_libs.sh <dir> <rounds> name1 name2 ...   (dir holds name.so files; run under gpurun)
d=$1; n=$2; shift 2
for r in $(seq $n); do for v in "$@"; do
  cp $d/$v.so rrin_b200/librrin_b200.so
  echo "== $v"; python tools/step_table.py 2>&1 | grep -E "^forward"; BATCH=4 python tools/step_table.py 2>&1 | grep -E "^forward"
done; done
