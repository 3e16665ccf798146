echo "== cfg1 single slice 30k"; BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.01 10 1 0 3 2>&1 | head -20
echo "== cfg5 single slice 1M"; BF_PROFILE=1 timeout 200 python tools/prof_batch.py 0.01 -1 1 0 3 100e6 1280 720 2>&1 | head -20
