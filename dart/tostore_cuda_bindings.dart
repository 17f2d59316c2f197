// tostore_cuda_bindings.dart — dart:ffi binding of libtostore_cuda.so
// (include/tostore_cuda.h). Drop into lib/src/handler/ of tocreator/tostore.
//
// Style follows the reference's only FFI user, lib/src/handler/system_ffi_helper.dart:
// a static DynamicLibrary, lookupFunction<Native, Dart>('name'), caller-allocated
// buffers from `calloc` freed in try/finally, int32 status (0 = success), and no
// exception crossing the public helper (errors degrade to an empty result, like
// every missing-precondition case of VectorIndexManager.vectorSearch,
// lib/src/core/vector_index_manager.dart:485-508).
//
// NOT compiled in this repository (no Dart SDK in the build image); it is kept
// small and one-to-one with the header so it can be reviewed by eye. The Python
// twin that IS exercised by the test-suite is tostore_b200/_native.py.
//
// Keep off the web build with the reference's conditional-import pattern
// (lib/src/handler/platform_handler.dart:1-2): `if (dart.library.io)`.

import 'dart:ffi';
import 'dart:typed_data';

import 'dart:convert';

import 'package:ffi/ffi.dart';

/// Mirrors `tsc_index_desc`.
final class TscIndexDesc extends Struct {
  @Uint32()
  external int structSize;
  @Uint32()
  external int dims;
  @Uint8()
  external int metric; // VectorDistanceMetric.index (l2, innerProduct, cosine)
  @Uint8()
  external int srcPrecision; // VectorPrecision.index (float64, float32, int8)
  @Uint8()
  external int devDtype; // 0 f32, 1 bf16, 2 f16
  @Uint8()
  external int reserved0;
  @Int32()
  external int deviceId;
  @Uint64()
  external int capacityRows;
  @Uint64()
  external int firstNodeId;
  @Uint32()
  external int kMax;
  @Uint32()
  external int nqMax;
  @Uint32()
  external int nDevices; // 0 / 1: one shard on deviceId; 2..8: a GROUP over deviceIds
  @Array(8)
  external Array<Int32> deviceIds;
  @Uint32()
  external int reserved1;
}

/// Mirrors `tsc_where_op`: one step of the postfix condition program
/// (leaf operators restate ConditionRecordMatcher._evaluateOperator,
/// lib/src/handler/value_matcher.dart:570-612).
final class TscWhereOp extends Struct {
  @Uint8()
  external int kind; // 0 leaf, 1 AND, 2 OR
  @Uint8()
  external int op; // 0 = 1 != 2 > 3 >= 4 < 5 <= 6 BETWEEN 7 IN 8 NOT IN 9 IS NULL 10 IS NOT NULL
  @Uint16()
  external int n;
  @Uint32()
  external int columnId;
  @Int64()
  external int iLo;
  @Int64()
  external int iHi;
  @Double()
  external double fLo;
  @Double()
  external double fHi;
  @Uint32()
  external int argsOffset;
  @Uint32()
  external int reserved;
}

typedef _CreateN = Int32 Function(Pointer<TscIndexDesc>, Pointer<Uint64>);
typedef _CreateD = int Function(Pointer<TscIndexDesc>, Pointer<Uint64>);
typedef _HandleN = Int32 Function(Uint64);
typedef _HandleD = int Function(int);
typedef _AppendRowsN = Int32 Function(Uint64, Uint64, Pointer<Void>, Uint64);
typedef _AppendRowsD = int Function(int, int, Pointer<Void>, int);
typedef _AppendPagesN = Int32 Function(
    Uint64, Uint64, Pointer<Uint8>, Uint64, Uint32, Uint64);
typedef _AppendPagesD = int Function(
    int, int, Pointer<Uint8>, int, int, int);
typedef _SetDeletedN = Int32 Function(Uint64, Pointer<Uint64>, Uint64, Uint8);
typedef _SetDeletedD = int Function(int, Pointer<Uint64>, int, int);
typedef _SearchN = Int32 Function(Uint64, Pointer<Float>, Uint32, Uint32, Double,
    Pointer<Int64>, Pointer<Double>, Pointer<Uint32>);
typedef _SearchD = int Function(int, Pointer<Float>, int, int, double,
    Pointer<Int64>, Pointer<Double>, Pointer<Uint32>);
typedef _SubmitN = Int32 Function(Uint64, Pointer<Float>, Uint32, Uint32, Double,
    Pointer<Int64>, Pointer<Double>, Pointer<Uint32>, Pointer<Uint64>);
typedef _SubmitD = int Function(int, Pointer<Float>, int, int, double,
    Pointer<Int64>, Pointer<Double>, Pointer<Uint32>, Pointer<Uint64>);
typedef _PollN = Int32 Function(Uint64, Pointer<Int32>);
typedef _PollD = int Function(int, Pointer<Int32>);
typedef _ColCreateN = Int32 Function(Uint64, Uint32, Uint8);
typedef _ColCreateD = int Function(int, int, int);
typedef _ColAppendN = Int32 Function(
    Uint64, Uint32, Uint64, Pointer<Void>, Pointer<Uint8>, Uint64);
typedef _ColAppendD = int Function(int, int, int, Pointer<Void>, Pointer<Uint8>, int);
typedef _FilterWhereN = Int32 Function(
    Uint64, Pointer<TscWhereOp>, Uint32, Pointer<Void>, Uint32, Pointer<Uint64>);
typedef _FilterWhereD = int Function(
    int, Pointer<TscWhereOp>, int, Pointer<Void>, int, Pointer<Uint64>);
typedef _SetPksN = Int32 Function(Uint64, Uint64, Pointer<Uint8>, Pointer<Uint64>, Uint64);
typedef _SetPksD = int Function(int, int, Pointer<Uint8>, Pointer<Uint64>, int);
typedef _SearchPkN = Int32 Function(
    Uint64, Pointer<Double>, Uint64, Uint32, Double, Pointer<Int64>, Pointer<Double>,
    Pointer<Double>, Pointer<Uint8>, Uint64, Pointer<Uint64>, Pointer<Uint32>);
typedef _SearchPkD = int Function(
    int, Pointer<Double>, int, int, double, Pointer<Int64>, Pointer<Double>,
    Pointer<Double>, Pointer<Uint8>, int, Pointer<Uint64>, Pointer<Uint32>);
typedef _SearchBatchN = Int32 Function(Uint64, Pointer<Double>, Uint64, Uint32, Uint32, Double,
    Pointer<Int64>, Pointer<Double>, Pointer<Double>, Pointer<Uint32>);
typedef _SearchBatchD = int Function(int, Pointer<Double>, int, int, int, double,
    Pointer<Int64>, Pointer<Double>, Pointer<Double>, Pointer<Uint32>);
typedef _LoadNghN = Int32 Function(Uint64, Pointer<Utf8>, Uint32, Pointer<Void>);
typedef _LoadNghD = int Function(int, Pointer<Utf8>, int, Pointer<Void>);
typedef _LastErrorN = Pointer<Utf8> Function();
typedef _LastErrorD = Pointer<Utf8> Function();

/// One (nodeId, distance) pair — same shape as `NghSearchResult`
/// (lib/src/core/ngh_graph_engine.dart:26-40).
class TscHit {
  final int nodeId;
  final double distance;
  const TscHit(this.nodeId, this.distance);
}

class TostoreCuda {
  static DynamicLibrary? _lib;
  static bool _failed = false;

  static DynamicLibrary? _open() {
    if (_lib != null || _failed) return _lib;
    try {
      _lib = DynamicLibrary.open('libtostore_cuda.so');
    } catch (_) {
      _failed = true; // no GPU library: callers keep using NghGraphEngine.search
    }
    return _lib;
  }

  static bool get isAvailable {
    final lib = _open();
    if (lib == null) return false;
    try {
      final count =
          lib.lookupFunction<Int32 Function(), int Function()>('tsc_device_count');
      return count() > 0;
    } catch (_) {
      return false;
    }
  }

  static String lastError() {
    final lib = _open();
    if (lib == null) return 'libtostore_cuda.so not found';
    return lib
        .lookupFunction<_LastErrorN, _LastErrorD>('tsc_last_error')()
        .toDartString();
  }

  /// tsc_index_create. Returns 0 on failure. With `deviceIds` (2..8 CUDA devices) the
  /// handle is a GROUP: the column is row-range sharded over those GPUs inside the library
  /// and every call below (append, delete, filter, search ...) works on it unchanged — one
  /// `search` call scans all shards and merges their exact top-k over NVLink. This is how
  /// the single-process Dart host uses all GPUs of a box.
  static int createIndex({
    required int dims,
    required int metricIndex,
    required int precisionIndex,
    required int capacityRows,
    int devDtype = 0,
    int deviceId = 0,
    List<int>? deviceIds,
    int firstNodeId = 0,
    int kMax = 128,
    int nqMax = 64,
  }) {
    final lib = _open();
    if (lib == null) return 0;
    final create = lib.lookupFunction<_CreateN, _CreateD>('tsc_index_create');
    final desc = calloc<TscIndexDesc>();
    final out = calloc<Uint64>();
    try {
      desc.ref
        ..structSize = sizeOf<TscIndexDesc>()
        ..dims = dims
        ..metric = metricIndex
        ..srcPrecision = precisionIndex
        ..devDtype = devDtype
        ..deviceId = deviceId
        ..capacityRows = capacityRows
        ..firstNodeId = firstNodeId
        ..kMax = kMax
        ..nqMax = nqMax
        ..nDevices = deviceIds == null ? 0 : deviceIds.length;
      for (var i = 0; deviceIds != null && i < deviceIds.length && i < 8; i++) {
        desc.ref.deviceIds[i] = deviceIds[i];
      }
      return create(desc, out) == 0 ? out.value : 0;
    } finally {
      calloc.free(desc);
      calloc.free(out);
    }
  }

  /// tsc_device_count: GPUs visible to the process (0 when there is none / on error).
  static int deviceCount() {
    final lib = _open();
    if (lib == null) return 0;
    final n = lib.lookupFunction<Int32 Function(), int Function()>('tsc_device_count')();
    return n < 0 ? 0 : n;
  }

  /// tsc_search_flags: verdict of the exactness certificate for the last search, per query:
  /// 0 exact, 2 = more than 4096 rows tie with the k-th neighbour within the key error bound
  /// (best-effort result). Empty list on error.
  static List<int> searchFlags(int handle, int nq) {
    final lib = _open();
    if (lib == null || nq <= 0) return const [];
    final fn = lib.lookupFunction<Int32 Function(Uint64, Uint32, Pointer<Uint32>),
        int Function(int, int, Pointer<Uint32>)>('tsc_search_flags');
    final buf = calloc<Uint32>(nq);
    try {
      if (fn(handle, nq, buf) != 0) return const [];
      return List<int>.generate(nq, (i) => buf[i]);
    } finally {
      calloc.free(buf);
    }
  }

  static void destroyIndex(int handle) {
    _open()?.lookupFunction<_HandleN, _HandleD>('tsc_index_destroy')(handle);
  }

  /// Flush-time hook: the Float32List rows VectorIndexManager.writeChanges just
  /// handed to NghGraphEngine.insertBatch (vector_index_manager.dart:378-387).
  static bool appendRows(int handle, int firstNodeId, Float32List rows, int nRows) {
    final lib = _open();
    if (lib == null || nRows == 0) return lib != null;
    final fn = lib.lookupFunction<_AppendRowsN, _AppendRowsD>('tsc_index_append_rows');
    final buf = calloc<Float>(rows.length);
    try {
      buf.asTypedList(rows.length).setAll(0, rows);
      return fn(handle, firstNodeId, buf.cast<Void>(), nRows) == 0;
    } finally {
      calloc.free(buf);
    }
  }

  /// Cold start: raw bytes of consecutive rawvec pages exactly as
  /// StorageInterface.readAsBytesAt returned them (ngh_partition_manager.dart:262-264).
  static bool appendPages(int handle, int firstLogicalPage, Uint8List pages,
      int pageSize, int liveRows) {
    final lib = _open();
    if (lib == null) return false;
    final fn = lib.lookupFunction<_AppendPagesN, _AppendPagesD>('tsc_index_append_pages');
    final buf = calloc<Uint8>(pages.length);
    try {
      buf.asTypedList(pages.length).setAll(0, pages);
      return fn(handle, firstLogicalPage, buf, pages.length ~/ pageSize, pageSize,
              liveRows) ==
          0;
    } finally {
      calloc.free(buf);
    }
  }

  /// deleteBatch hook (vector_index_manager.dart:429-434).
  static bool setDeleted(int handle, List<int> nodeIds, {bool deleted = true}) {
    final lib = _open();
    if (lib == null || nodeIds.isEmpty) return lib != null;
    final fn = lib.lookupFunction<_SetDeletedN, _SetDeletedD>('tsc_index_set_deleted');
    final buf = calloc<Uint64>(nodeIds.length);
    try {
      for (var i = 0; i < nodeIds.length; i++) {
        buf[i] = nodeIds[i];
      }
      return fn(handle, buf, nodeIds.length, deleted ? 1 : 0) == 0;
    } finally {
      calloc.free(buf);
    }
  }

  /// Blocking search: replaces the body of the `_graphEngine.search` call at
  /// vector_index_manager.dart:538-548. `query` is the already padded (and, for
  /// cosine, normalised) Float32List. Returns [] on any error.
  static List<TscHit> search(int handle, Float32List query, int topK,
      {double? distanceThreshold}) {
    final lib = _open();
    if (lib == null || topK <= 0) return const [];
    final fn = lib.lookupFunction<_SearchN, _SearchD>('tsc_search');
    final q = calloc<Float>(query.length);
    final ids = calloc<Int64>(topK);
    final dist = calloc<Double>(topK);
    final count = calloc<Uint32>();
    try {
      q.asTypedList(query.length).setAll(0, query);
      final rc = fn(handle, q, 1, topK, distanceThreshold ?? double.nan, ids, dist, count);
      if (rc != 0) return const [];
      return [for (var i = 0; i < count.value; i++) TscHit(ids[i], dist[i])];
    } finally {
      calloc.free(q);
      calloc.free(ids);
      calloc.free(dist);
      calloc.free(count);
    }
  }

  /// Non-blocking variant: submit, then `await Future.delayed(Duration.zero)`
  /// between polls, the way YieldController keeps the isolate responsive.
  static Future<List<TscHit>> searchAsync(int handle, Float32List query, int topK,
      {double? distanceThreshold}) async {
    final lib = _open();
    if (lib == null || topK <= 0) return const [];
    final submit = lib.lookupFunction<_SubmitN, _SubmitD>('tsc_search_submit');
    final poll = lib.lookupFunction<_PollN, _PollD>('tsc_search_poll');
    final q = calloc<Float>(query.length);
    final ids = calloc<Int64>(topK);
    final dist = calloc<Double>(topK);
    final count = calloc<Uint32>();
    final ticket = calloc<Uint64>();
    final done = calloc<Int32>();
    try {
      q.asTypedList(query.length).setAll(0, query);
      if (submit(handle, q, 1, topK, distanceThreshold ?? double.nan, ids, dist, count,
              ticket) !=
          0) {
        return const [];
      }
      while (true) {
        if (poll(ticket.value, done) != 0) return const [];
        if (done.value != 0) break;
        await Future<void>.delayed(Duration.zero);
      }
      return [for (var i = 0; i < count.value; i++) TscHit(ids[i], dist[i])];
    } finally {
      calloc.free(q);
      calloc.free(ids);
      calloc.free(dist);
      calloc.free(count);
      calloc.free(ticket);
      calloc.free(done);
    }
  }
  /// Attribute column for the GPU WHERE prefilter: `colType` 0 = integer, 1 = double, 2 = text
  /// (DataType.integer / DataType.double fields of the table).
  static bool columnCreate(int handle, int columnId, int colType) {
    final lib = _open();
    if (lib == null) return false;
    return lib.lookupFunction<_ColCreateN, _ColCreateD>('tsc_index_column_create')(
            handle, columnId, colType) ==
        0;
  }

  /// Flush-time hook next to appendRows: the field's values for node ids
  /// [firstNodeId, firstNodeId + values.length); `null` entries become NULL.
  static bool columnAppendInt(int handle, int columnId, int firstNodeId, List<int?> values) {
    final lib = _open();
    if (lib == null || values.isEmpty) return lib != null;
    final fn = lib.lookupFunction<_ColAppendN, _ColAppendD>('tsc_index_column_append');
    final v = calloc<Int64>(values.length);
    final nulls = calloc<Uint8>(values.length);
    try {
      for (var i = 0; i < values.length; i++) {
        v[i] = values[i] ?? 0;
        nulls[i] = values[i] == null ? 1 : 0;
      }
      return fn(handle, columnId, firstNodeId, v.cast(), nulls, values.length) == 0;
    } finally {
      calloc.free(v);
      calloc.free(nulls);
    }
  }

  static bool columnAppendDouble(int handle, int columnId, int firstNodeId, List<double?> values) {
    final lib = _open();
    if (lib == null || values.isEmpty) return lib != null;
    final fn = lib.lookupFunction<_ColAppendN, _ColAppendD>('tsc_index_column_append');
    final v = calloc<Double>(values.length);
    final nulls = calloc<Uint8>(values.length);
    try {
      for (var i = 0; i < values.length; i++) {
        v[i] = values[i] ?? 0.0;
        nulls[i] = values[i] == null ? 1 : 0;
      }
      return fn(handle, columnId, firstNodeId, v.cast(), nulls, values.length) == 0;
    } finally {
      calloc.free(v);
      calloc.free(nulls);
    }
  }

  /// Text attribute column (`colType` 2, DataType.text): values travel as their UTF-16 code
  /// units (`String.codeUnits`), so the GPU's comparisons are `String.compareTo` and LIKE is
  /// `ValueMatcher.matchesLike` (handler/value_matcher.dart:211-240, :318-331). Pass the values
  /// the table stores, i.e. after `FieldSchema.convertValue` (trim()). `null` -> NULL.
  static bool columnAppendText(int handle, int columnId, int firstNodeId, List<String?> values) {
    final lib = _open();
    if (lib == null || values.isEmpty) return lib != null;
    final fn = lib.lookupFunction<
        Int32 Function(Uint64, Uint32, Uint64, Pointer<Uint16>, Pointer<Uint64>, Pointer<Uint8>, Uint64),
        int Function(int, int, int, Pointer<Uint16>, Pointer<Uint64>, Pointer<Uint8>,
            int)>('tsc_index_column_append_text');
    final total = values.fold<int>(0, (a, s) => a + (s?.length ?? 0));
    final units = calloc<Uint16>(total == 0 ? 1 : total);
    final offs = calloc<Uint64>(values.length + 1);
    final nulls = calloc<Uint8>(values.length);
    try {
      var o = 0;
      for (var i = 0; i < values.length; i++) {
        offs[i] = o;
        final s = values[i];
        nulls[i] = s == null ? 1 : 0;
        if (s != null) {
          units.asTypedList(total == 0 ? 1 : total).setRange(o, o + s.length, s.codeUnits);
          o += s.length;
        }
      }
      offs[values.length] = o;
      return fn(handle, columnId, firstNodeId, units, offs, nulls, values.length) == 0;
    } finally {
      calloc.free(units);
      calloc.free(offs);
      calloc.free(nulls);
    }
  }

  /// `filterWhere` for conditions with leaves on text columns: `texts` is the operand pool
  /// (already normalised: `condition.normalize` trims text operands); a text leaf names its
  /// operand by index in `i_lo` (`i_hi`: BETWEEN end), its IN list holds indices.
  static int filterWhereText(int handle, Pointer<TscWhereOp> ops, int nOps, Pointer<Void> inArgs,
      int nInArgs, List<String> texts) {
    final lib = _open();
    if (lib == null) return -1;
    final fn = lib.lookupFunction<
        Int32 Function(Uint64, Pointer<TscWhereOp>, Uint32, Pointer<Void>, Uint32, Pointer<Uint16>,
            Pointer<Uint64>, Uint32, Pointer<Uint64>),
        int Function(int, Pointer<TscWhereOp>, int, Pointer<Void>, int, Pointer<Uint16>,
            Pointer<Uint64>, int, Pointer<Uint64>)>('tsc_index_filter_where_text');
    final total = texts.fold<int>(0, (a, s) => a + s.length);
    final units = calloc<Uint16>(total == 0 ? 1 : total);
    final offs = calloc<Uint64>(texts.length + 1);
    final matched = calloc<Uint64>();
    try {
      var o = 0;
      for (var i = 0; i < texts.length; i++) {
        offs[i] = o;
        units.asTypedList(total == 0 ? 1 : total).setRange(o, o + texts[i].length, texts[i].codeUnits);
        o += texts[i].length;
      }
      offs[texts.length] = o;
      return fn(handle, ops, nOps, inArgs, nInArgs, units, offs, texts.length, matched) == 0
          ? matched.value
          : -1;
    } finally {
      calloc.free(units);
      calloc.free(offs);
      calloc.free(matched);
    }
  }

  /// Evaluate a compiled condition (postfix `ops`, filled by the caller from
  /// `QueryCondition.build()` after `normalize`) on the GPU and install it as the
  /// prefilter of the following searches. Returns the number of matching rows, -1 on error.
  static int filterWhere(int handle, Pointer<TscWhereOp> ops, int nOps, Pointer<Void> inArgs,
      int nInArgs) {
    final lib = _open();
    if (lib == null) return -1;
    final fn = lib.lookupFunction<_FilterWhereN, _FilterWhereD>('tsc_index_filter_where');
    final matched = calloc<Uint64>();
    try {
      return fn(handle, ops, nOps, inArgs, nInArgs, matched) == 0 ? matched.value : -1;
    } finally {
      calloc.free(matched);
    }
  }

  /// `__nid2pk` deltas of `_writeNodeMappings` (vector_index_manager.dart:1276-1293):
  /// keys for node ids [firstNodeId, ...); an empty string is the tombstone mapping.
  static bool setPrimaryKeys(int handle, int firstNodeId, List<String> pks) {
    final lib = _open();
    if (lib == null || pks.isEmpty) return lib != null;
    final fn = lib.lookupFunction<_SetPksN, _SetPksD>('tsc_index_set_primary_keys');
    final enc = [for (final p in pks) utf8.encode(p)];
    final total = enc.fold<int>(0, (a, b) => a + b.length);
    final bytes = calloc<Uint8>(total == 0 ? 1 : total);
    final offs = calloc<Uint64>(pks.length + 1);
    try {
      var o = 0;
      for (var i = 0; i < enc.length; i++) {
        offs[i] = o;
        bytes.asTypedList(total == 0 ? 1 : total).setRange(o, o + enc[i].length, enc[i]);
        o += enc[i].length;
      }
      offs[pks.length] = o;
      return fn(handle, firstNodeId, bytes, offs, pks.length) == 0;
    } finally {
      calloc.free(bytes);
      calloc.free(offs);
    }
  }

  /// WHERE prefilter from a set of primary keys (the result of any `db.query(...)`): the rows
  /// whose key is in `pks` stay searchable for the following searches. Replaces the
  /// `__pk2nid` lookups (vector_index_manager.dart:1350-1363) + `setFilter`. Returns the number
  /// of rows selected, -1 on error.
  static int filterPrimaryKeys(int handle, List<String> pks) {
    final lib = _open();
    if (lib == null) return -1;
    final fn = lib.lookupFunction<
        Int32 Function(Uint64, Pointer<Uint8>, Pointer<Uint64>, Uint64, Pointer<Uint64>),
        int Function(int, Pointer<Uint8>, Pointer<Uint64>, int,
            Pointer<Uint64>)>('tsc_index_filter_primary_keys');
    final enc = [for (final p in pks) utf8.encode(p)];
    final total = enc.fold<int>(0, (a, b) => a + b.length);
    final bytes = calloc<Uint8>(total == 0 ? 1 : total);
    final offs = calloc<Uint64>(pks.length + 1);
    final matched = calloc<Uint64>();
    try {
      var o = 0;
      for (var i = 0; i < enc.length; i++) {
        offs[i] = o;
        bytes.asTypedList(total == 0 ? 1 : total).setRange(o, o + enc[i].length, enc[i]);
        o += enc[i].length;
      }
      offs[pks.length] = o;
      return fn(handle, bytes, offs, pks.length, matched) == 0 ? matched.value : -1;
    } finally {
      calloc.free(bytes);
      calloc.free(offs);
      calloc.free(matched);
    }
  }

  /// Whole `VectorIndexManager.vectorSearch` body after the precondition checks
  /// (vector_index_manager.dart:514-588) in one call: query prep, exact search, nodeId -> PK,
  /// score. Returns (primaryKey, distance, score) triples, ascending distance; [] on error.
  static List<(String, double, double)> vectorSearchPk(
      int handle, List<double> queryVector, int topK,
      {double? distanceThreshold, int pkCapacity = 65536}) {
    final lib = _open();
    if (lib == null || topK <= 0) return const [];
    final fn = lib.lookupFunction<_SearchPkN, _SearchPkD>('tsc_vector_search_pk');
    final q = calloc<Double>(queryVector.isEmpty ? 1 : queryVector.length);
    final ids = calloc<Int64>(topK);
    final dist = calloc<Double>(topK);
    final score = calloc<Double>(topK);
    final pkBytes = calloc<Uint8>(pkCapacity);
    final offs = calloc<Uint64>(topK + 1);
    final count = calloc<Uint32>();
    try {
      for (var i = 0; i < queryVector.length; i++) {
        q[i] = queryVector[i];
      }
      final rc = fn(handle, q, queryVector.length, topK, distanceThreshold ?? double.nan, ids,
          dist, score, pkBytes, pkCapacity, offs, count);
      if (rc != 0) return const [];
      final raw = pkBytes.asTypedList(pkCapacity);
      return [
        for (var i = 0; i < count.value; i++)
          (utf8.decode(raw.sublist(offs[i], offs[i + 1])), dist[i], score[i])
      ];
    } finally {
      calloc.free(q);
      calloc.free(ids);
      calloc.free(dist);
      calloc.free(score);
      calloc.free(pkBytes);
      calloc.free(offs);
      calloc.free(count);
    }
  }
  /// Cold start: stream `<indexDir>/ngh/{rawvec,graph}/dir_k/p<n>.ngh` into the GPU index
  /// (`tsc_index_load_ngh`: reader thread + pinned double buffer inside the library;
  /// pages are validated and decoded on the GPU). `indexDir` is the directory that holds
  /// `ngh/meta.json` (core/path_manager.dart:317-324).
  static bool loadNgh(int handle, String indexDir, {bool tombstones = true}) {
    final lib = _open();
    if (lib == null) return false;
    final fn = lib.lookupFunction<_LoadNghN, _LoadNghD>('tsc_index_load_ngh');
    final dir = indexDir.toNativeUtf8(allocator: calloc);
    try {
      return fn(handle, dir, tombstones ? 1 : 0, nullptr) == 0;
    } finally {
      calloc.free(dir);
    }
  }
  /// Batch form of the search (additive; `ToStore.vectorSearch` stays single-query):
  /// `queries` are equally long fp64 vectors; returns per query the (nodeId, distance, score)
  /// triples in ascending distance order. [] on error.
  static List<List<(int, double, double)>> vectorSearchBatch(
      int handle, List<List<double>> queries, int topK,
      {double? distanceThreshold}) {
    final lib = _open();
    if (lib == null || topK <= 0 || queries.isEmpty) return const [];
    final fn = lib.lookupFunction<_SearchBatchN, _SearchBatchD>('tsc_vector_search_batch');
    final nq = queries.length, len = queries.first.length;
    final q = calloc<Double>(nq * (len == 0 ? 1 : len));
    final ids = calloc<Int64>(nq * topK);
    final dist = calloc<Double>(nq * topK);
    final score = calloc<Double>(nq * topK);
    final counts = calloc<Uint32>(nq);
    try {
      for (var i = 0; i < nq; i++) {
        for (var c = 0; c < len && c < queries[i].length; c++) {
          q[i * len + c] = queries[i][c];
        }
      }
      if (fn(handle, q, len, nq, topK, distanceThreshold ?? double.nan, ids, dist, score, counts) != 0) {
        return const [];
      }
      return [
        for (var i = 0; i < nq; i++)
          [for (var j = 0; j < counts[i]; j++) (ids[i * topK + j], dist[i * topK + j], score[i * topK + j])]
      ];
    } finally {
      calloc.free(q);
      calloc.free(ids);
      calloc.free(dist);
      calloc.free(score);
      calloc.free(counts);
    }
  }
}
