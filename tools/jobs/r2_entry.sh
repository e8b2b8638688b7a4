python tools/entry_profile.py 2e6
python tools/entry_profile.py 1e7 4e6
python tools/entry_profile.py 1e7 1e6
nproc; free -g | head -2; df -h /tmp /dev/shm | tail -2
