for v in b256k b1m b4m b16m; do echo "== $v"; PTB_LIB_PATH=pytestbeam_b200/_lib/var_$v.so python tools/entry_profile.py 2e7 2>&1 | grep "files in\|whole entry"; done
timeout 300 python tools/entry_time.py 1e8 2 2>&1 | grep -v INFO | tail -3
