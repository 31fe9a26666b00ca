#!/bin/bash
# round-2 GPU call BB: smoke() with the bit-for-bit comparison
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
