# A/B of compile-time knobs on the training step: bash profiles/ab_train.sh "NRMS_ATTN_BWD_HC20=15" "NRMS_ATTN_BWD_HC20=5"
for d in "$@"; do
  NRMS_DEFINES="$d" python newsrecommendationsystem_b200/csrc/build.py --force > /dev/null 2>&1
  echo "== $d"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file /tmp/tr.csv python profiles/train_step_probe.py 3 > /tmp/tr.log 2>&1
  python profiles/summarize_launches.py /tmp/tr.csv | head -8
done
