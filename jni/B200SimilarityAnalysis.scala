// SOURCE ONLY -- not compiled or tested in this image (no JVM / scalac).  See INTEGRATION.md.
// Drop-in for org.apache.mahout.math.cf.SimilarityAnalysis as called at
// /root/reference/src/main/scala/URAlgorithm.scala:323-329 (cooccurrencesIDSs) and :343-346 (crossOccurrenceDownsampled).
// The change in the reference is the import at URAlgorithm.scala:26:
//   import com.actionml.b200.{ DownsamplableCrossOccurrenceDataset, B200SimilarityAnalysis => SimilarityAnalysis }
package com.actionml.b200

import java.nio.{ ByteBuffer, ByteOrder }

import scala.collection.JavaConverters._

import org.apache.mahout.math.SequentialAccessSparseVector
import org.apache.mahout.math.indexeddataset.IndexedDataset
import org.apache.mahout.sparkbindings._
import org.apache.mahout.sparkbindings.indexeddataset.IndexedDatasetSpark

/** same fields as org.apache.mahout.math.cf.DownsamplableCrossOccurrenceDataset as built at URAlgorithm.scala:336-340 */
case class DownsamplableCrossOccurrenceDataset(iD: IndexedDataset, maxElementsPerRow: Int = 500,
  maxInterestingElements: Int = 50, minLLROpt: Option[Double] = None, parOpts: Option[Any] = None)

/** driver-side CSR of one IndexedDataset in pinned, native-order direct buffers (what cco_csr_t points at) */
final class Csr(val rowPtr: ByteBuffer, val colIdx: ByteBuffer, val nnz: Long)

object Csr {
  /** rows = (userIdx, item indices of that user).  Users without a row keep an empty range; indices may be unsorted and
    * may repeat (setQuick semantics, Preparator.scala:201-208): the library canonicalises on the device. */
  def pack(nRows: Long, rows: Array[(Int, Array[Int])], alloc: Long => ByteBuffer): Csr = {
    val deg = new Array[Int](nRows.toInt)
    var i = 0
    while (i < rows.length) { deg(rows(i)._1) += rows(i)._2.length; i += 1 }
    val rp = alloc(8L * (nRows + 1)).order(ByteOrder.nativeOrder)
    var acc = 0L
    var r = 0
    while (r < nRows) { rp.putLong(8 * r, acc); acc += deg(r); r += 1 }
    rp.putLong(8 * nRows.toInt, acc)
    val ci = alloc(math.max(4L * acc, 4L)).order(ByteOrder.nativeOrder)
    val cursor = new Array[Long](nRows.toInt)
    i = 0
    while (i < rows.length) {
      val (u, items) = rows(i)
      var q = rp.getLong(8 * u) + cursor(u)
      var j = 0
      while (j < items.length) { ci.putInt((4 * q).toInt, items(j)); q += 1; j += 1 }
      cursor(u) += items.length
      i += 1
    }
    new Csr(rp, ci, acc)
  }
}

object B200SimilarityAnalysis {
  System.loadLibrary("cco_b200_jni") // jni/cco_jni.c, links libcco_b200.so
  sys.addShutdownHook(shutdown())

  @native private def hostAlloc(bytes: Long): ByteBuffer // pinned (cco_host_alloc)
  @native private def hostFree(buf: ByteBuffer): Unit
  @native private def train(rowPtr: Array[ByteBuffer], colIdx: Array[ByteBuffer], nRows: Long, nCols: Array[Int],
    m: Array[Int], k: Array[Int], hasMinLlr: Array[Boolean], minLlr: Array[Double], seed: Int, flags: Int): Long // cco_result_t*
  @native private def resultNumMatrices(res: Long): Int
  @native private def resultMatrix(res: Long, i: Int): Array[ByteBuffer] // row_ptr(int64), col(int32), llr(f64) views
  @native private def resultFree(res: Long): Unit
  @native private def shutdown(): Unit

  /** same signature as Mahout's SimilarityAnalysis.crossOccurrenceDownsampled (URAlgorithm.scala:343) */
  def crossOccurrenceDownsampled(datasets: List[DownsamplableCrossOccurrenceDataset], randomSeed: Int = 0xdeadbeef): List[IndexedDataset] = {
    val a = datasets.head.iD.asInstanceOf[IndexedDatasetSpark]
    implicit val sc = a.matrix.context.asInstanceOf[SparkDistributedContext].sc
    val nRows = a.matrix.nrow
    // driver-side CSR: collect each DRM's (userIdx, Vector) rows; values are all 1.0 so only indices travel
    val csr = datasets.map { d =>
      val rows = d.iD.matrix.rdd.map { case (r, v) => r -> v.nonZeroes.iterator().asScala.map(_.index).toArray }.collect()
      Csr.pack(nRows, rows, hostAlloc)
    }
    val res = try train(csr.map(_.rowPtr).toArray, csr.map(_.colIdx).toArray, nRows, datasets.map(_.iD.matrix.ncol).toArray,
      datasets.map(_.maxElementsPerRow).toArray, datasets.map(_.maxInterestingElements).toArray,
      datasets.map(_.minLLROpt.isDefined).toArray, datasets.map(_.minLLROpt.getOrElse(0.0)).toArray, randomSeed, 0)
    finally csr.foreach { c => hostFree(c.rowPtr); hostFree(c.colIdx) } // the library never keeps host pointers
    try datasets.zipWithIndex.map {
      case (d, i) =>
        val Array(rp, ci, llr) = resultMatrix(res, i).map(_.order(ByteOrder.nativeOrder))
        val nItemsA = a.matrix.ncol
        val rows = (0 until nItemsA).map { r =>
          val (s, e) = (rp.getLong(8 * r).toInt, rp.getLong(8 * (r + 1)).toInt)
          val v = new SequentialAccessSparseVector(d.iD.matrix.ncol, e - s)
          var q = s; while (q < e) { v.setQuick(ci.getInt(4 * q), llr.getDouble(8 * q)); q += 1 }
          r -> (v: org.apache.mahout.math.Vector)
        }
        val drm = drmWrap[Int](sc.parallelize(rows), nrow = nItemsA, ncol = d.iD.matrix.ncol)
        a.create(drm, a.columnIDs, d.iD.columnIDs) // what Mahout returns: rows = A's items, cols = B's items
    } finally resultFree(res)
  }

  /** same signature as SimilarityAnalysis.cooccurrencesIDSs (URAlgorithm.scala:323) */
  def cooccurrencesIDSs(indexedDatasets: Array[IndexedDataset], randomSeed: Int = 0xdeadbeef,
    maxInterestingItemsPerThing: Int = 50, maxNumInteractions: Int = 500): List[IndexedDataset] =
    crossOccurrenceDownsampled(indexedDatasets.toList.map(
      DownsamplableCrossOccurrenceDataset(_, maxNumInteractions, maxInterestingItemsPerThing, None)), randomSeed)
}
