#!/bin/bash
# Rebuild the UNMODIFIED reference (alibaba/graph-learn at /root/reference) offline and install it into baseline/_ref.
# baseline/_ref is git-ignored, so it has to be rebuilt whenever the working tree is re-created from the repository
# (~12 minutes on 8 cores).  Everything happens in a scratch copy under /tmp: /root/reference is read-only.
#   bash baseline/build_reference.sh
# Steps: gRPC 1.38.1 + protobuf + abseil + c-ares + re2 from third_party/grpc, glog, then the
# graphlearn core (no KNN / hiactor / vineyard), `make python` -> wheel -> pip install --no-index --no-deps --target.
# The only source edit is in the scratch copy of the THIRD-PARTY abseil (std::max(SIGSTKSZ, ...) does not compile with
# glibc >= 2.34); `-include cstdint` covers headers that gcc 13 no longer pulls in transitively.
set -euo pipefail
HERE=$(dirname "$(realpath "$0")")
W=${REF_BUILD_DIR:-/tmp/refbuild}
J=${JOBS:-$(nproc)}
rm -rf "$W" && cp -r /root/reference "$W" && chmod -R u+w "$W"

cd "$W/third_party/grpc/grpc"
sed -i 's/std::max(SIGSTKSZ, 65536)/std::max<size_t>(SIGSTKSZ, 65536)/' third_party/abseil-cpp/absl/debugging/failure_signal_handler.cc
mkdir -p cmake/build && cd cmake/build
cmake -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_CXX_FLAGS="-fPIC -include cstdint -Wno-error" -DCMAKE_C_FLAGS="-fPIC" -DCMAKE_BUILD_TYPE=Release \
  -DgRPC_INSTALL=ON -DCMAKE_INSTALL_PREFIX="$W/third_party/grpc/build" -DgRPC_BUILD_TESTS=OFF -DgRPC_SSL_PROVIDER=package -DgRPC_ZLIB_PROVIDER=package \
  -DgRPC_BUILD_CSHARP_EXT=OFF -DgRPC_BUILD_GRPC_CSHARP_PLUGIN=OFF -DgRPC_BUILD_GRPC_NODE_PLUGIN=OFF -DgRPC_BUILD_GRPC_OBJECTIVE_C_PLUGIN=OFF \
  -DgRPC_BUILD_GRPC_PHP_PLUGIN=OFF -DgRPC_BUILD_GRPC_PYTHON_PLUGIN=OFF -DgRPC_BUILD_GRPC_RUBY_PLUGIN=OFF ../.. > "$W/grpc_cmake.log" 2>&1
make -j"$J" > "$W/grpc_make.log" 2>&1
make install > "$W/grpc_install.log" 2>&1
cp -r "$W/third_party/grpc/grpc/third_party/abseil-cpp/absl" "$W/third_party/grpc/build/include"
cp -r "$W/third_party/grpc/grpc/third_party/cares/cares" "$W/third_party/grpc/build/include"

cd "$W/third_party/glog/glog" && mkdir -p build && cd build
cmake -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_CXX_FLAGS="-fPIC" -DCMAKE_INSTALL_PREFIX="$W/third_party/glog/build" -DBUILD_SHARED_LIBS=OFF \
  -DBUILD_TESTING=OFF -DWITH_GFLAGS=OFF .. > "$W/glog_cmake.log" 2>&1
make -j"$J" > "$W/glog_make.log" 2>&1 && make install > "$W/glog_install.log" 2>&1

# setup.py / CMakeLists ask git for a revision string: give the scratch copy a repository
cd "$W" && git init -q . && git -c user.email=ref@local -c user.name=ref add -A graphlearn/setup > /dev/null && \
  git -c user.email=ref@local -c user.name=ref commit -qm "reference snapshot" > /dev/null
cd "$W/graphlearn" && mkdir -p cmake-build && cd cmake-build
cmake -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DKNN=OFF -DWITH_HIACTOR=OFF -DWITH_VINEYARD=OFF -DTESTING=OFF -DGL_CXX_DIALECT=c++17 \
  -DCMAKE_CXX_FLAGS="-include cstdint" -DGL_PYTHON_BIN="$(which python)" .. > "$W/gl_cmake.log" 2>&1
make -j"$J" graphlearn_shared > "$W/gl_make.log" 2>&1
make python > "$W/gl_python.log" 2>&1

rm -rf "$HERE/_ref"
python -m pip install --no-index --no-deps --target "$HERE/_ref" "$W"/graphlearn/dist/graph_learn-1.2.0-*.whl
echo "installed: $(ls "$HERE/_ref")"
