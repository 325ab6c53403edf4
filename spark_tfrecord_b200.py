"""Import shim: the package directory is named ``spark-tfrecord_b200/`` (the reference repo's name
plus the target), which is not a Python identifier.  This module gives it the importable name
``spark_tfrecord_b200`` by pointing ``__path__`` at that directory, so
``import spark_tfrecord_b200.sqltypes`` etc. resolve to files inside it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "spark-tfrecord_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
