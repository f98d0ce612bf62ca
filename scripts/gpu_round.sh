#!/bin/bash
bash scripts/collect_profiles.sh r02
