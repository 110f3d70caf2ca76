"""Parses the cc65 .dbg output to extract symbol addresses (reference
transcoder/symbol_table.py)."""

from typing import Dict, TextIO


class SymbolTable:
    """Parse cc65 debug file to extract symbol table."""

    def __init__(self, debugfile: str = None):
        self.debugfile = debugfile

    def parse(self, iostream: TextIO = None) -> Dict:
        """name -> {key: value} for every ``sym`` line of the debug file."""
        if not iostream:
            iostream = open(self.debugfile, "r")
        syms = {}
        with iostream as f:
            for line in f.read().split("\n"):
                if not line.startswith("sym"):
                    continue
                sym = dict(kv.split("=") for kv in line.split()[1].split(","))
                syms[sym["name"]] = sym
        return syms
