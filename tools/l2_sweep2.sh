CF=""
for cap in 12 16 20 24 30 38; do for w in 0 2 4; do for h in 1 2; do CF="$CF;1,$cap,$w,$h"; done; done; done
CF="${CF:1};1,38,0,3;1,24,0,3;2,16,2,2;2,24,2,2;2,38,2,2;2,38,0,2"
python tools/l2_sweep.py 300000 "$CF" 1.8 2>&1 | grep "^n="
python tools/l2_sweep.py 100000 "$CF" 1.2 2>&1 | grep "^n="
