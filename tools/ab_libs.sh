# same-box A/B of several builds: bash tools/ab_libs.sh <dir> <rounds> <grep pattern> name1 name2 ...   (dir holds name.so; run under gpurun)
d=$1; n=$2; pat=$3; shift 3
for r in $(seq $n); do for v in "$@"; do
  cp $d/$v.so rrin_b200/librrin_b200.so
  echo "== $v"; python tools/step_table.py 2>&1 | grep -E "^forward|$pat"; BATCH=4 python tools/step_table.py 2>&1 | grep -E "^forward|$pat"
done; done
