for n in 20000 100000; do
for cap in 24 32 38 48; do
  echo "== n=$n rm_cap=$cap panels=off"; SEMGATE_RM_CAP_MB=$cap timeout 200 python tools/size_sweep.py $n 1,2 3 2>&1 | grep "^n=" | tail -n +2 | awk '{print $1,$2,$3,$4,$5,$6}' | sort | uniq -c | sort -rn | head -4
done; done
for cap in 24 38; do for pm in 16 24; do
  echo "== n=300000 rm_cap=$cap panel_mb=$pm"; SEMGATE_RM_CAP_MB=$cap SEMGATE_PANEL_MB=$pm timeout 200 python tools/size_sweep.py 300000 1,2 4 2>&1 | grep "^n=" | awk '{print $1,$2,$3,$4}' | tail -4
done; done
for cap in 24 38; do
  echo "== n=300000 rm_cap=$cap panels=off"; SEMGATE_RM_CAP_MB=$cap timeout 200 python tools/size_sweep.py 300000 1,2 4 2>&1 | grep "^n=" | awk '{print $1,$2,$3,$4}' | tail -4
done
