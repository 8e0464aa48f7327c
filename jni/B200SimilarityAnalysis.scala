// SOURCE ONLY -- not compiled or tested in this image (no JVM).  See INTEGRATION.md.
// Drop-in for org.apache.mahout.math.cf.SimilarityAnalysis as called at URAlgorithm.scala:323-329, 343-346.
package com.actionml.b200

import java.nio.{ByteBuffer, ByteOrder}
import org.apache.mahout.math.SequentialAccessSparseVector
import org.apache.mahout.math.indexeddataset.IndexedDataset
import org.apache.mahout.sparkbindings._
import org.apache.mahout.sparkbindings.indexeddataset.IndexedDatasetSpark

case class DownsamplableCrossOccurrenceDataset(iD: IndexedDataset, maxElementsPerRow: Int = 500,
  maxInterestingElements: Int = 50, minLLROpt: Option[Double] = None)

object B200SimilarityAnalysis {
  System.loadLibrary("cco_b200_jni")                       // jni/cco_jni.c, links libcco_b200.so

  @native private def train(rowPtr: Array[ByteBuffer], colIdx: Array[ByteBuffer], nRows: Long, nCols: Array[Int],
    m: Array[Int], k: Array[Int], hasMinLlr: Array[Boolean], minLlr: Array[Double], seed: Int): Long   // -> cco_result_t*
  @native private def resultMatrix(res: Long, i: Int): Array[ByteBuffer]   // row_ptr(int64), col(int32), llr(f64)
  @native private def resultFree(res: Long): Unit

  /** same signature as Mahout's SimilarityAnalysis.crossOccurrenceDownsampled (URAlgorithm.scala:343) */
  def crossOccurrenceDownsampled(datasets: List[DownsamplableCrossOccurrenceDataset], randomSeed: Int = 0xdeadbeef)
      : List[IndexedDataset] = {
    val a = datasets.head.iD.asInstanceOf[IndexedDatasetSpark]
    implicit val sc = a.matrix.context.asInstanceOf[SparkDistributedContext].sc
    val nRows = a.matrix.nrow
    // driver-side CSR: collect each DRM's (userIdx, Vector) rows; values are all 1.0 so only indices travel
    val csr = datasets.map { d =>
      val rows = d.iD.matrix.rdd.map { case (r, v) => r -> v.nonZeroes.iterator().asScala.map(_.index).toArray }.collect()
      Csr.pack(nRows, rows)                               // direct, native-order ByteBuffers (row_ptr int64, col_idx int32)
    }
    val res = train(csr.map(_.rowPtr).toArray, csr.map(_.colIdx).toArray, nRows, datasets.map(_.iD.matrix.ncol).toArray,
      datasets.map(_.maxElementsPerRow).toArray, datasets.map(_.maxInterestingElements).toArray,
      datasets.map(_.minLLROpt.isDefined).toArray, datasets.map(_.minLLROpt.getOrElse(0.0)).toArray, randomSeed)
    try datasets.zipWithIndex.map { case (d, i) =>
      val Array(rp, ci, llr) = resultMatrix(res, i).map(_.order(ByteOrder.nativeOrder))
      val nItemsA = a.matrix.ncol
      val rows = (0 until nItemsA).map { r =>
        val (s, e) = (rp.getLong(8 * r).toInt, rp.getLong(8 * (r + 1)).toInt)
        val v = new SequentialAccessSparseVector(d.iD.matrix.ncol, e - s)
        var q = s; while (q < e) { v.setQuick(ci.getInt(4 * q), llr.getDouble(8 * q)); q += 1 }
        r -> (v: org.apache.mahout.math.Vector)
      }
      val drm = drmWrap[Int](sc.parallelize(rows), nrow = nItemsA, ncol = d.iD.matrix.ncol)
      a.create(drm, a.columnIDs, d.iD.columnIDs)        // what Mahout returns: rows = A's items, cols = B's items
    } finally resultFree(res)
  }

  /** same signature as SimilarityAnalysis.cooccurrencesIDSs (URAlgorithm.scala:323) */
  def cooccurrencesIDSs(indexedDatasets: Array[IndexedDataset], randomSeed: Int = 0xdeadbeef,
      maxInterestingItemsPerThing: Int = 50, maxNumInteractions: Int = 500): List[IndexedDataset] =
    crossOccurrenceDownsampled(indexedDatasets.toList.map(
      DownsamplableCrossOccurrenceDataset(_, maxNumInteractions, maxInterestingItemsPerThing, None)), randomSeed)
}
