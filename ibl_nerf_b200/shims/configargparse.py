"""Minimal stand-in for `configargparse` (not installed in this image), covering exactly what the reference's
config_parser.py uses: ArgumentParser(default_config_files=...), add_argument(..., is_config_file=True),
parse_args() with a list, None or ONE STRING (config_parser.py:8), config files with `key = value`,
`key=True/False` and bare `key` lines for store_true flags, `[a, b]` lists for action="append", `#` comments;
later files and the command line override earlier ones."""
import argparse
import shlex
import sys


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *args, default_config_files=None, **kwargs):
        super().__init__(*args, **kwargs)
        self._default_config_files = list(default_config_files or [])
        self._config_dests = []

    def add_argument(self, *args, is_config_file=False, **kwargs):
        action = super().add_argument(*args, **kwargs)
        if is_config_file:
            self._config_dests.append(action.dest)
        return action

    def _file_tokens(self, path):
        toks = []
        with open(path) as f:
            for line in f:
                line = line.split("#", 1)[0].strip()
                if not line:
                    continue
                if "=" in line:
                    key, val = [t.strip() for t in line.split("=", 1)]
                else:
                    key, val = line, None
                opt = "--" + key
                action = self._option_string_actions.get(opt)
                if action is None:
                    raise SystemExit("%s: unknown option %r" % (path, key))
                if isinstance(action, argparse._StoreTrueAction):
                    if val is None or val.lower() in ("true", "1", "yes"):
                        toks.append(opt)
                elif val is not None and val.startswith("[") and val.endswith("]"):
                    for item in val[1:-1].split(","):
                        if item.strip():
                            toks += [opt, item.strip()]
                else:
                    toks += [opt, val if val is not None else ""]
        return toks

    def parse_known_args(self, args=None, namespace=None):
        if args is None:
            args = sys.argv[1:]
        elif isinstance(args, str):
            args = shlex.split(args)
        args = list(args)
        # config files named on the command line
        pre = argparse.ArgumentParser(add_help=False)
        for dest in self._config_dests:
            pre.add_argument("--" + dest, default=None)
        known, _ = pre.parse_known_args(args)
        files = list(self._default_config_files) + [getattr(known, d) for d in self._config_dests if getattr(known, d)]
        toks = []
        for f in files:
            toks += self._file_tokens(f)
        return super().parse_known_args(toks + args, namespace)


ArgParser = ArgumentParser
