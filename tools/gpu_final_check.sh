#!/bin/bash
timeout 100 python tools/unet_error.py 2>&1 | tail -1
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
