#!/bin/bash
cd "$GRAFT_REPO_ROOT"
DC_RADIUS_FILL=direct timeout 300 python tools/prof_radius.py 2>&1 | tail -5
timeout 300 python tools/prof_radius.py 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q -k "radius or nn_ or golden or feature or dropin" 2>&1 | tail -4
