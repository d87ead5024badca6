#!/bin/bash
# Copy the development tree build/work (edited while a gpurun call may still be queueing on the repo snapshot) into the repo.
set -e
SRC=/root/repo/build/work
DST=/root/repo
cd "$SRC"
tar --exclude="*.so" --exclude=__pycache__ --exclude=.pytest_cache --exclude="./oracle/_build" --exclude="./build" \
    --exclude="./gpurun_out" --exclude="./.hypothesis" -cf /tmp/azb_sync.tar .
cd "$DST"
tar xf /tmp/azb_sync.tar
rm -f /tmp/azb_sync.tar
