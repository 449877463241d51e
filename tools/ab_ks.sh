# usage: bash tools/ab_ks.sh variant ... : tools/ks_time.py once per library variant ("default" = the shipped build)
for v in "$@"; do
  if [ "$v" = default ]; then python tools/ks_time.py --tag default $KS_ARGS
  else ABC_B200_LIB=$PWD/abc_b200/lib/libabc_b200_$v.so python tools/ks_time.py --tag $v $KS_ARGS; fi
done
