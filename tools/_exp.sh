cd /root/repo
run() { echo "== $*"; env $1 python tools/quick_time.py 16384 $2 $3 2>&1 | grep -v WARN | tail -1 | sed 's/units=16384 //;s/hop2=None //;s/nprb=273 linear filter: //;s/est\/s=[0-9]* //;s/units\/s=[0-9]* //'; }
B=SRSCE_LIBRARY=/root/repo/tools/_lib_base.so
for r in 1 2; do
  run "$B" 13 comb4
  run "X=0" 13 comb4
  run "$B" 3,10 comb4
  run "X=0" 3,10 comb4
  run "$B" 3 comb2
  run "X=0" 3 comb2
done
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -2
