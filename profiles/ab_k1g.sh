# A/B of compile-time K1g knobs on the evaluate probe: bash profiles/ab_k1g.sh "K1G_V_RELOAD=0" "K1G_V_RELOAD=1" ...
for d in "$@"; do
  NRMS_DEFINES="$d" python newsrecommendationsystem_b200/csrc/build.py --force > /dev/null 2>&1
  echo "== $d"; timeout 200 python profiles/k1g_probe.py 2>&1 | grep -A1 "kernel total" | sed -n 4,5p
done
