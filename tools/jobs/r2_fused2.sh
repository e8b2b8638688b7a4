python tools/stage_time.py
python tools/time_variants.py tools/stage_time.py nocopy nolb noflush
