O=gpurun_out/r2af; mkdir -p $O
L=$PWD/gstpeaq_b200
for v in "" _NOPF; do PEAQ_B200_LIBRARY=$L/libpeaq_b200$v.so python scripts/time_modes.py advanced 2>&1 | grep libpeaq; done > $O/log.txt
cat $O/log.txt
python scripts/compare_builds.py $L/libpeaq_b200_NOPF.so $L/libpeaq_b200.so 2>&1 | tail -3
