#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
