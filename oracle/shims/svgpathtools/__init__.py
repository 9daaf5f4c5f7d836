"""Import-time stand-in for svgpathtools (the reference's Datasets/svg_parser.py imports it at module level; the
SVG -> graph preprocessing is out of scope and never runs here)."""


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)
    return type(name, (), {})
