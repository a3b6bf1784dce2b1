# same-box A/B of a variant library (oblas_b200/liboblas_b200_exp.so) against the in-tree one:
# tools/tc_classes.py per-kernel-class times, lazy and full outer-panel flows, alternating
cd $GRAFT_REPO_ROOT
V=$PWD/oblas_b200/liboblas_b200_exp.so
for i in 1 2; do
  echo "== base lazy";  python tools/tc_classes.py | tail -2 | head -1
  echo "== exp  lazy";  OB2_LIB=$V python tools/tc_classes.py | tail -2 | head -1
done
echo "== base full"; LAZY=0 python tools/tc_classes.py | tail -2 | head -1
echo "== exp  full"; LAZY=0 OB2_LIB=$V python tools/tc_classes.py | tail -2 | head -1
