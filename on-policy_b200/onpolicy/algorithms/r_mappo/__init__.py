from pkgutil import extend_path as _extend_path

__path__ = _extend_path(__path__, __name__)   # fall through to a reference checkout for modules not provided here
