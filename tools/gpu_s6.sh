V=abvariants
echo "--- pretest on"; python tools/detect_probe.py 2>&1 | tail -5
echo "--- pretest compiled out"; CPM_B200_LIB=$V/nopre/libcpm_b200.so CPM_HOST_LIB=$V/nopre/libcpm_host.so python tools/detect_probe.py 2>&1 | tail -5
