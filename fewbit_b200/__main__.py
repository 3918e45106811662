"""Command line of the package, after the reference's `fewbit` command (fewbit/cli.py:130-181):

    python -m fewbit_b200 quantize [-o tables.npz] NOBITS SPEC     build and save a few-bit table
    python -m fewbit_b200 version
    python -m fewbit_b200 help
"""
import argparse
import sys

from . import __version__, quantize


def main(argv=None):
    parser = argparse.ArgumentParser(prog='python -m fewbit_b200', description=__doc__.split('\n\n')[0])
    sub = parser.add_subparsers(dest='command')
    sub.add_parser('help', add_help=False, help='Show this message and exit.')
    quantize.add_arguments(sub.add_parser('quantize', help='Build and save few-bit approximation.'))
    sub.add_parser('version', add_help=False, help='Show version information.')
    args = parser.parse_args(argv)
    if args.command == 'quantize':
        quantize.run(args)
    elif args.command == 'version':
        print(f'fewbit_b200 version {__version__}')
    elif args.command == 'help':
        parser.print_help()
    else:
        parser.print_usage()


if __name__ == '__main__':
    main(sys.argv[1:])
