for d in "K1G_EX2_FMA=0" "K1G_EX2_FMA=2" "K1G_EX2_FMA=4"; do
  NRMS_DEFINES="$d" python newsrecommendationsystem_b200/csrc/build.py --force > /dev/null 2>&1
  echo "== $d"; timeout 200 python profiles/k1g_probe.py 2>&1 | tail -n 2 | head -n 1
done
