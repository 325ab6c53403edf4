"""spark_tfrecord_b200 -- B200-native TFRecord decode/encode behind the spark-tfrecord API.

Layout (hot path only):
  csrc/        sm_100a CUDA kernels + the C ABI (libtfrgpu.so, include/tfrgpu.h)
  _native.py   ctypes binding of the C ABI (fails loudly if the library or a GPU is missing)
  sqltypes.py  StructType/StructField/... mirror of the Spark SQL types the reference takes
  io.py        host-side mirror of the reference interface: TFRecordFileReader.readFile,
               TFRecordOutputWriter, TFRecordDeserializer / TFRecordSerializer, DefaultSource
"""
__version__ = "0.1.0"
