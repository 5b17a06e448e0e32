#!/usr/bin/env python
"""python run_recbole.py -m FOCF -d ml-100k -c my.yaml   (the reference's run_recbole.py:16-26 over recbole_fairrec_b200)"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--model", "-m", type=str, default="FOCF", help="name of models")
    parser.add_argument("--dataset", "-d", type=str, default="ml-100k", help="name of datasets")
    parser.add_argument("--config_files", "-c", type=str, default=None, help="config files (space separated)")
    parser.add_argument("--saved", action="store_true", help="check-point the best model and evaluate the test split with it")
    args, rest = parser.parse_known_args()
    from recbole_fairrec_b200.quick_start import run_recbole
    files = args.config_files.strip().split(" ") if args.config_files else None
    out = run_recbole(model=args.model, dataset=args.dataset, config_file_list=files, saved=args.saved, argv=rest)
    print(json.dumps({k: (dict(v) if hasattr(v, "items") else v) for k, v in out.items()}, indent=1, default=float))
