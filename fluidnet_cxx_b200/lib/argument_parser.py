"""`lib.SmartFormatter` (reference: pytorch/lib/argument_parser.py): help strings starting with 'R|' keep
their own line breaks; used by the drivers' argparse set-up (plume.py:36, rayleighTaylor.py:36)."""
import argparse


class SmartFormatter(argparse.HelpFormatter):
    def _split_lines(self, text, width):
        if text.startswith('R|'):
            return text[2:].splitlines()
        return super()._split_lines(text, width)
