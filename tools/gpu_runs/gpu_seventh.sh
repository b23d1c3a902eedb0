python tools/check_status.py
