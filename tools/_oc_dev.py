import os, sys
sys.path.insert(0, '/root/repo')
from ndtpso_slam_b200 import capi
capi._build.LIB_PATH = '/root/repo/tools/_build/lib_dev.so'
exec(open('/root/repo/tools/onecall_chunks.py').read())
